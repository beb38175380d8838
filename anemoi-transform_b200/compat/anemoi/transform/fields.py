from anemoi_transform_b200.fields import *  # noqa: F401,F403
from anemoi_transform_b200.fields import (  # noqa: F401
    FieldSelection,
    NewDataField,
    NewLatLonField,
    NewMetadataField,
    WrappedField,
    new_empty_fieldlist,
    new_field_from_latitudes_longitudes,
    new_field_from_numpy,
    new_field_with_metadata,
    new_fieldlist_from_list,
)
