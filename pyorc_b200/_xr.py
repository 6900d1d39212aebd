"""Tiny stand-in for the two xarray types the PIV path touches, used ONLY when xarray is not importable
(this build image has no xarray; a pyorc installation always has it and then the real classes are used).

Just enough surface for the binding and the parity tests to read like the reference's:
``len(da)``, ``da[0].shape``, ``da[a:b]``, ``da.values``, ``da.load()``, ``da.time``, ``ds["v_x"].values``,
``ds.mean(dim="time", keep_attrs=True)``, ``concat([...], dim="time")``.
"""

from __future__ import annotations

import numpy as np


class DataArray:
    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(self.values.ndim))
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.attrs = dict(attrs or {})
        self.name = name
        self.encoding = {}

    # -- basic protocol -----------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    def __len__(self):
        return self.values.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __getattr__(self, item):
        coords = self.__dict__.get("coords", {})
        if item in coords:
            dims = (item,) if coords[item].ndim == 1 else None
            return DataArray(coords[item], dims=dims, coords={item: coords[item]} if dims else None)
        raise AttributeError(item)

    def __getitem__(self, key):
        if isinstance(key, str):
            return getattr(self, key)
        vals = self.values[key]
        if isinstance(key, (int, np.integer)):
            dims = self.dims[1:]
            coords = {k: v for k, v in self.coords.items() if k != self.dims[0]}
        else:
            dims = self.dims
            coords = dict(self.coords)
            if self.dims[0] in coords:
                coords[self.dims[0]] = coords[self.dims[0]][key]
        return DataArray(vals, dims=dims, coords=coords, attrs=self.attrs, name=self.name)

    def load(self):
        return self

    def copy(self, deep=True):
        return DataArray(self.values.copy() if deep else self.values, self.dims, dict(self.coords), dict(self.attrs), self.name)

    def diff(self, dim):
        ax = self.dims.index(dim)
        coords = dict(self.coords)
        if dim in coords:
            coords[dim] = coords[dim][1:]
        return DataArray(np.diff(self.values, axis=ax), self.dims, coords, self.attrs, self.name)

    def mean(self, dim=None, keep_attrs=False, skipna=True):
        ax = None if dim is None else self.dims.index(dim)
        fn = np.nanmean if (skipna and np.issubdtype(self.values.dtype, np.floating)) else np.mean
        with np.errstate(all="ignore"):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                vals = fn(self.values, axis=ax)
        dims = tuple(d for d in self.dims if d != dim) if dim is not None else ()
        coords = {k: v for k, v in self.coords.items() if k != dim}
        return DataArray(vals, dims, coords, self.attrs if keep_attrs else None, self.name)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.attrs = dict(attrs or {})
        self.data_vars = {}
        for k, v in (data_vars or {}).items():
            if isinstance(v, DataArray):
                self.data_vars[k] = v
            else:
                dims, data = v[0], v[1]
                self.data_vars[k] = DataArray(data, dims, {d: self.coords[d] for d in dims if d in self.coords}, name=k)

    def __getitem__(self, k):
        if k in self.data_vars:
            return self.data_vars[k]
        if k in self.coords:
            return DataArray(self.coords[k], (k,), {k: self.coords[k]})
        raise KeyError(k)

    def __contains__(self, k):
        return k in self.data_vars or k in self.coords

    def __getattr__(self, item):
        d = self.__dict__
        if item in d.get("data_vars", {}) or item in d.get("coords", {}):
            return self[item]
        raise AttributeError(item)

    def keys(self):
        return self.data_vars.keys()

    def mean(self, dim=None, keep_attrs=False):
        out = Dataset({}, {k: v for k, v in self.coords.items() if k != dim}, self.attrs if keep_attrs else None)
        for k, v in self.data_vars.items():
            out.data_vars[k] = v.mean(dim=dim, keep_attrs=keep_attrs)
        return out


def concat(objs, dim):
    first = objs[0]
    coords = dict(first.coords)
    if dim in coords:
        coords[dim] = np.concatenate([o.coords[dim] for o in objs])
    out = Dataset({}, coords, first.attrs)
    for k, v in first.data_vars.items():
        ax = v.dims.index(dim)
        out.data_vars[k] = DataArray(np.concatenate([o.data_vars[k].values for o in objs], axis=ax), v.dims,
                                     {d: coords[d] for d in v.dims if d in coords}, v.attrs, k)
    return out
