"""numpy stand-in for the three fastremap functions the reference's Python package calls (crackle/codec.py:74,765;
crackle/operations.py:94,104,235,245,497,576).  fastremap is a third-party dependency of the reference that is not
installed in this image.  Test infrastructure only."""
import numpy as np


def unique(labels, return_counts=False, return_index=False, return_inverse=False):
    return np.unique(np.asarray(labels), return_counts=return_counts, return_index=return_index, return_inverse=return_inverse)


def fit_dtype(dtype, value, exotics=False):
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.floating):
        return dtype
    value = int(value)
    if np.issubdtype(dtype, np.signedinteger) or value < 0:
        for dt in (np.int8, np.int16, np.int32, np.int64):
            if np.iinfo(dt).min <= value <= np.iinfo(dt).max:
                return np.dtype(dt)
        raise ValueError(f"Unable to find a compatible dtype for {dtype} that can fit {value}")
    for dt in (np.uint8, np.uint16, np.uint32, np.uint64):
        if value <= np.iinfo(dt).max:
            return np.dtype(dt)
    raise ValueError(f"Unable to find a compatible dtype for {dtype} that can fit {value}")


def remap(arr, table, preserve_missing_labels=False, in_place=False):
    arr = np.asarray(arr)
    src = arr.reshape(-1).tolist()
    if not in_place and len(table):          # like fastremap: a copy is widened when the new values do not fit
        wide = fit_dtype(arr.dtype, max(int(v) for v in table.values()))
        if np.dtype(wide).itemsize > arr.dtype.itemsize:
            arr = arr.astype(wide)
    a = arr if in_place else np.array(arr, copy=True)
    flat = a.reshape(-1)
    for i, v in enumerate(src):
        if v in table:
            flat[i] = table[v]
        elif not preserve_missing_labels:
            raise KeyError(f"{v} was not in the remap table.")
    return a
