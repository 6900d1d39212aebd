"""Generate tests/golden/ngwerere_masks.npz from the reference's own example files (run in the build container, where
/root/reference exists; the fixture travels, the reference does not).

examples/03_Plotting_and_masking_velocimetry_results.ipynb (cells 3, 10, 16) opens `ngwerere/ngwerere_piv.nc`, applies

    ds_mask2.velocimetry.mask.corr(inplace=True)
    ds_mask2.velocimetry.mask.minmax(inplace=True)
    ds_mask2.velocimetry.mask.rolling(inplace=True)
    ds_mask2.velocimetry.mask.outliers(inplace=True)
    ds_mask2.velocimetry.mask.variance(inplace=True)
    ds_mask2.velocimetry.mask.angle(angle_tolerance=0.5*np.pi)          # not in place: returns a mask, changes nothing
    ds_mask2.velocimetry.mask.count(inplace=True)
    ds_mask2.velocimetry.mask.window_mean(wdw=2, inplace=True, tolerance=0.5, reduce_time=True)
    ds_mask2.velocimetry.set_encoding(); ds_mask2.to_netcdf("ngwerere_masked.nc")

and the reference ships BOTH files: the input and the output of its own mask stack (xarray + pyorc/api/mask.py), int16 CF-packed
with the encoding of pyorc/const.py:80-83.  They are netCDF4 = HDF5; there is no HDF5 reader in this image, tests/golden/h5min.py
parses just enough of the format.  The fixture holds the packed input fields (v_x, v_y, corr; s2n is not used by these masks and holds
int16 wrap-around values), the packing attributes as found in the file, and which values survive in the output.
(`ngwerere_piv.nc` itself was written by an older pyorc - 25 px windows, stride 13, OpenPIV-style axes - so its VALUES pin nothing
about today's PIV engine; as the input of the mask stack that does not matter.)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import h5min  # noqa: E402

REF = "/root/reference/examples/ngwerere"


def load(fn):
    h = h5min.H5(os.path.join(REF, fn))
    names = h5min.dataset_names(h)
    out = {k: h.read(names[k]) for k in ("v_x", "v_y", "corr", "s2n", "time", "y", "x")}
    return out, h


def main():
    piv, h = load("ngwerere_piv.nc")
    msk, _ = load("ngwerere_masked.nc")
    scale = h5min.dense_attributes(h, "scale_factor")
    fills = [f for f in h5min.dense_attributes(h, "_FillValue") if np.asarray(f).dtype == np.int16]
    assert len(scale) == 4 and all(float(np.asarray(s).reshape(-1)[0]) == 0.01 for s in scale), scale
    assert len(fills) == 4 and all(int(np.asarray(f).reshape(-1)[0]) == -9999 for f in fills), fills
    fill = -9999
    kept = msk["v_x"] != fill
    for k in ("v_x", "v_y", "corr", "s2n"):
        assert piv[k].dtype == np.int16 and piv[k].shape == (125, 59, 66)
        assert np.array_equal(msk[k] != fill, kept), k                        # one mask for all four variables
        assert np.array_equal(msk[k][kept], piv[k][kept]), k                  # survivors are unchanged: decode -> encode is the identity
        assert not (kept & (piv[k] == fill)).any()
    np.savez_compressed(
        os.path.join(HERE, "ngwerere_masks.npz"),
        v_x=piv["v_x"], v_y=piv["v_y"], corr=piv["corr"], kept_bits=np.packbits(kept), shape=np.array(kept.shape),
        scale_factor=np.float64(0.01), fill_value=np.int16(fill), time=piv["time"], y=piv["y"], x=piv["x"],
        s2n_min_max=np.array([piv["s2n"].min(), piv["s2n"].max()]),
        provenance=np.array("reference examples/ngwerere/ngwerere_piv.nc (input) and ngwerere_masked.nc (output of notebook 03, cells 10 + 16)"),
    )
    print("kept", kept.mean(), "of", kept.size, "values; input NaN fraction", (piv["v_x"] == fill).mean())


if __name__ == "__main__":
    main()
