"""Weight initialisers (reference: nn/initializers.py:4-25).  Host-side NumPy: drawn once per
layer, then uploaded.  Note the reference's `*_uniform` draws U[0, 1) (all-positive weights),
which is kept."""
import numpy as np


def xavier_normal(in_num, out_num):
    return np.random.normal(size=(in_num, out_num)) / np.sqrt(in_num)


def xavier_uniform(in_num, out_num):
    return np.random.uniform(size=(in_num, out_num)) / np.sqrt(in_num)


def kaiming_normal(in_num, out_num):
    return np.random.normal(size=(in_num, out_num)) / np.sqrt(in_num / 2)


def kaiming_uniform(in_num, out_num):
    return np.random.uniform(size=(in_num, out_num)) / np.sqrt(in_num / 2)
