"""Stand-in for healpy (not installed): lets the reference's tabular support module import.
HEALPix grids themselves are not available offline."""


def _unavailable(*args, **kwargs):
    raise ImportError("healpy is not installed in this environment")


nside2npix = pix2ang = _unavailable
