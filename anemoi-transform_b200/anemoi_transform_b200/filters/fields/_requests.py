"""`patch_data_request` helpers shared by the filters that trade one param for another
(reference: the per-filter `patch_data_request` methods, e.g. lnsp_to_sp.py:69-98)."""

from __future__ import annotations

from typing import Any


def swap_param(data_request: dict[str, Any], a: str, b: str, both_error: str) -> dict[str, Any]:
    """Ask for `b` where the request names `a` and vice versa; naming both is an error."""
    param = data_request.get("param")
    if param is None:
        return data_request
    listed = param if isinstance(param, list) else [param]
    if a in listed and b in listed:
        raise ValueError(both_error)
    for have, want in ((a, b), (b, a)):
        if have in listed:
            data_request["param"].remove(have)
            data_request["param"].append(want)
            break
    return data_request


def replace_products(data_request: dict[str, Any], products: tuple[str, ...], source: str) -> dict[str, Any]:
    """Ask for `source` instead of the fields a filter derives from it."""
    param = data_request.get("param")
    if param is None:
        return data_request
    if any(p in param for p in products):
        data_request["param"] = [p for p in param if p not in products] + [source]
    return data_request


def rename_on_levels(data_request: Any, a: str, b: str, both_error: str) -> Any:
    """On pressure levels (levtype "pl" or an explicit levelist) ask for `b` in place of `a`, or
    for `a` in place of `b`, keeping the position in the list (orog_to_z.py:79-93)."""
    param = data_request.get("param")
    if param is None:
        return data_request
    listed = param if isinstance(param, list) else [param]
    if a in listed and b in listed:
        raise ValueError(both_error)
    on_levels = data_request.get("levtype", "") == "pl" or data_request.get("levelist", [])
    if on_levels:
        swap = {a: b} if a in listed else {b: a} if b in listed else {}
        if swap:
            data_request["param"] = [swap.get(p, p) for p in listed]
    return data_request
