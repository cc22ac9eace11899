"""Extract the plotted curves from the reference's committed MATLAB figures (results/*.fig are MAT-5 files holding the
handle-graphics tree) into tests/golden/reference_figs.json.  Runs in the build container only (/root/reference is absent
on the GPU box); the JSON is what travels.  These are the only numeric outputs the reference ships for the path
(no tests, seeds or golden vectors), so they are the oracle's one external anchor.
The figures predate the plot scripts at HEAD (errorVSadmmiters.fig holds 70 iterations, the script runs 100;
errorVSsnr.fig holds 3 SNR points, the script sweeps 11), so the comparison in tests/test_oracle.py is statistical.
usage: python tools/extract_fig_curves.py [/root/reference/results]"""
import json
import os
import sys

import numpy as np
import scipy.io as sio

FIGS = ("errorVSadmmiters", "errorVSsnr", "errorVSdelays", "errorVSspatialpaths", "errorVStraining_hbf", "errorVStraining_dbf", "errorVStraining", "errorVSsnr_angles")


def curves(node, out):
    if not (isinstance(node, np.ndarray) and node.dtype.names):
        return out
    for it in node.flat:
        names = it.dtype.names
        if "properties" in names:
            pr = it["properties"]
            if isinstance(pr, np.ndarray) and pr.dtype.names and "XData" in pr.dtype.names and "YData" in pr.dtype.names:
                for p in pr.flat:
                    name = str(np.squeeze(p["DisplayName"])) if "DisplayName" in p.dtype.names else ""
                    x = np.atleast_1d(np.squeeze(p["XData"])).astype(float)
                    y = np.atleast_1d(np.squeeze(p["YData"])).astype(float)
                    if x.size == y.size and x.size > 1:
                        out.append({"name": name, "x": x.tolist(), "y": y.tolist()})
        if "children" in names:
            curves(it["children"], out)
    return out


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/results"
    doc = {"source": "vlaxose/jstsp19 results/*.fig (handle-graphics XData/YData of every line object, in tree order)", "figures": {}}
    for f in FIGS:
        path = os.path.join(src, f + ".fig")
        if not os.path.exists(path):
            continue
        try:
            d = sio.loadmat(path, squeeze_me=False, struct_as_record=True)
        except NotImplementedError:           # MAT v7.3 (HDF5) figure: no reader in this image
            print(f, "skipped (MAT v7.3)")
            continue
        key = [k for k in d if k.startswith("hgS")][0]
        doc["figures"][f] = curves(d[key], [])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_figs.json")
    with open(dst, "w") as fh:
        json.dump(doc, fh, indent=0)
    for f, c in doc["figures"].items():
        print(f, [(k["name"], len(k["x"])) for k in c])


if __name__ == "__main__":
    main()
