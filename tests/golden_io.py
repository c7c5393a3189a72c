"""Loader for tests/golden/vectors/*.npz (written by tests/golden/make_golden_vectors.py)."""
import os

import numpy as np
import scipy.sparse as sp

VEC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors")


def load(name):
    z = np.load(os.path.join(VEC, name + ".npz"))
    out, groups = {}, {}
    for k in z.files:
        if "__" in k:
            g, f = k.split("__")
            groups.setdefault(g, {})[f] = z[k]
        else:
            out[k] = z[k]
    for g, f in groups.items():
        out[g] = sp.csc_matrix((f["data"], f["indices"], f["indptr"]), shape=tuple(f["shape"]))
    return out


def chol_inputs(g):
    L = {"perm": g["perm"], "L": g["Lpat"], "xsuper": g["xsuper"], "tmpsiz": float(g["tmpsiz"])}
    pars = {"abstol": float(g["pars"][0]), "canceltol": float(g["pars"][1]), "maxu": float(g["pars"][2])}
    return L, g["X"], pars, g["absd"]
