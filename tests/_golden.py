"""Loader for the committed golden fixtures (made by oracle/gen_golden.py from the unmodified reference)."""
import os

import numpy as np

from ppo_bihyb_b200.synthetic import HostGraph

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def pairs_of(z):
    out = []
    for p in range(len(z["n1"])):
        n1, n2 = int(z["n1"][p]), int(z["n2"][p])
        a1 = z["adj1"][z["adj1_off"][p]:z["adj1_off"][p + 1]].reshape(n1, n1)
        a2 = z["adj2"][z["adj2_off"][p]:z["adj2_off"][p + 1]].reshape(n2, n2)
        l1 = z["lab1"][z["lab1_off"][p]:z["lab1_off"][p + 1]]
        l2 = z["lab2"][z["lab2_off"][p]:z["lab2_off"][p + 1]]
        out.append((HostGraph(a1.copy(), l1.copy()), HostGraph(a2.copy(), l2.copy())))
    return out


def mats_of(z, key, pairs=None, off="off"):
    """Ragged (n1+1)x(n2+1) matrices stored flat under `key`."""
    out = []
    for p in range(len(z["n1"])):
        R, C = int(z["n1"][p]) + 1, int(z["n2"][p]) + 1
        out.append(z[key][z[off][p]:z[off][p + 1]].reshape(R, C))
    return out


def x_to_match(x):
    """Dense (n1+1)x(n2+1) 0/1 LSAPE solution -> (rowmatch [n1], colmatch [n2]) in the C-ABI's convention."""
    x = np.asarray(x)
    n1, n2 = x.shape[0] - 1, x.shape[1] - 1
    rowmatch = x[:n1].argmax(1).astype(np.int32) if n1 else np.zeros(0, np.int32)
    if n2:
        colmatch = np.where(x[n1, :n2] > 0, n1, x[:n1, :n2].argmax(0) if n1 else 0).astype(np.int32)
    else:
        colmatch = np.zeros(0, np.int32)
    return rowmatch, colmatch
