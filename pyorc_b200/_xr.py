"""Tiny stand-in for the two xarray types the PIV path touches, used ONLY when xarray is not importable
(this build image has no xarray; a pyorc installation always has it and then the real classes are used).

Just enough surface for the binding and the parity tests to read like the reference's:
``len(da)``, ``da[0].shape``, ``da[a:b]``, ``da.values``, ``da.load()``, ``da.time``, ``ds["v_x"].values``,
``ds.mean(dim="time", keep_attrs=True)``, ``concat([...], dim="time")`` - plus what the reference's own ``Frames.get_piv``
body needs around the engine call when it is run on these classes in the drop-in test (tests/test_reference_dropin.py):
accessor registration, attribute-style access to ``attrs``, ``diff``, ``assign_coords`` with 2-D coordinates.
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

    def __setattr__(self, key, value):
        object.__setattr__(self, key, value)
        owner = self.__dict__.get("_owner")
        if key == "attrs" and owner is not None:   # `ds[coord].attrs = {...}` (ORCBase.add_xy_coords)
            owner[0].coord_attrs[owner[1]] = value

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
        acc = _ACCESSORS["dataarray"].get(item)
        if acc is not None:
            return acc(self)
        attrs = self.__dict__.get("attrs", {})
        if item in attrs:   # xarray exposes attrs as attributes too (pyorc reads `da.camera_config`, `da.h_a`)
            return attrs[item]
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

    def sel(self, **indexers):
        """Label-based selection along 1-D coordinates (exact matches, like xarray's default)."""
        out = self
        for dim, labels in indexers.items():
            ax = out.dims.index(dim)
            labels = np.asarray(getattr(labels, "values", labels))
            pos = {v: i for i, v in enumerate(out.coords[dim].tolist())}
            idx = np.array([pos[v] for v in np.atleast_1d(labels).tolist()], dtype=np.int64)
            coords = dict(out.coords)
            coords[dim] = out.coords[dim][idx]
            out = DataArray(np.take(out.values, idx, axis=ax), out.dims, coords, out.attrs, out.name)
        return out

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


_ACCESSORS = {"dataarray": {}, "dataset": {}}


def register_dataarray_accessor(name):
    def deco(cls):
        _ACCESSORS["dataarray"][name] = cls
        return cls

    return deco


def register_dataset_accessor(name):
    def deco(cls):
        _ACCESSORS["dataset"][name] = cls
        return cls

    return deco


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.coord_dims = {k: (k,) for k in self.coords}     # dims of every coordinate (2-D ones come from assign_coords)
        self.coord_attrs = {}
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
            da = DataArray(self.coords[k], self.coord_dims.get(k, (k,)), {k: self.coords[k]} if self.coords[k].ndim == 1 else None, name=k)
            da.attrs = self.coord_attrs.setdefault(k, {})   # the same dict every time: `ds[k].attrs = {...}` is kept below
            da._owner = (self, k)
            return da
        raise KeyError(k)

    def assign_coords(self, coords):
        """``{name: (dims, values)}`` -> a new Dataset with the extra (possibly 2-D) coordinates."""
        out = Dataset({}, dict(self.coords), self.attrs)
        out.coord_dims, out.coord_attrs = dict(self.coord_dims), {k: dict(v) for k, v in self.coord_attrs.items()}
        out.data_vars = dict(self.data_vars)
        for k, v in coords.items():
            dims, vals = (v[0], v[1]) if isinstance(v, tuple) else ((k,), v)
            out.coords[k] = np.asarray(vals)
            out.coord_dims[k] = tuple(dims)
        return out

    def __contains__(self, k):
        return k in self.data_vars or k in self.coords

    def __getattr__(self, item):
        d = self.__dict__
        if item in d.get("data_vars", {}) or item in d.get("coords", {}):
            return self[item]
        acc = _ACCESSORS["dataset"].get(item)
        if acc is not None:
            return acc(self)
        if item in d.get("attrs", {}):
            return d["attrs"][item]
        raise AttributeError(item)

    def keys(self):
        return self.data_vars.keys()

    def mean(self, dim=None, keep_attrs=False):
        out = Dataset({}, {k: v for k, v in self.coords.items() if k != dim}, self.attrs if keep_attrs else None)
        out.coord_dims = {k: self.coord_dims.get(k, (k,)) for k in out.coords}
        out.coord_attrs = {k: dict(v) for k, v in self.coord_attrs.items() if k in out.coords}
        for k, v in self.data_vars.items():
            out.data_vars[k] = v.mean(dim=dim, keep_attrs=keep_attrs)
        return out


def concat(objs, dim):
    first = objs[0]
    coords = dict(first.coords)
    if dim in coords:
        coords[dim] = np.concatenate([o.coords[dim] for o in objs])
    out = Dataset({}, coords, first.attrs)
    out.coord_dims.update(first.coord_dims)
    for k, v in first.data_vars.items():
        ax = v.dims.index(dim)
        out.data_vars[k] = DataArray(np.concatenate([o.data_vars[k].values for o in objs], axis=ax), v.dims,
                                     {d: coords[d] for d in v.dims if d in coords}, v.attrs, k)
    return out
