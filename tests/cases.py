"""Seeded test cases shared by the golden-vector generator and the parity tests."""

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def si_model_kwargs():
    # tests/files/polymlp.in.phono3py.Si of the reference: 6 Gaussians (mu = 0..5) + the cutoff-only function
    return dict(n_type=1, cutoff=6.0, model_type=3, max_p=2, gtinv_order=3, gtinv_maxl=[4, 4],
                gaussian_params1=(1.0, 1.0, 1), gaussian_params2=(0.0, 5.0, 6))


def cfg2_model_kwargs(model_type=4):
    # BASELINE config 1/2/5 model: order 3, maxl [4,4], cutoff 6, 10 radial functions
    return dict(n_type=1, cutoff=6.0, model_type=model_type, max_p=2, gtinv_order=3, gtinv_maxl=[4, 4], n_gaussians=10)


def binary_model_kwargs():
    return dict(n_type=2, cutoff=5.0, model_type=4, max_p=2, gtinv_order=3, gtinv_maxl=[3, 2],
                pair_params=[[1.0, m] for m in np.linspace(0, 4, 5)] + [[0.0, 0.0]],
                pair_params_conditional={(0, 0): [0, 1, 2, 5], (0, 1): [1, 2, 3, 4, 5], (1, 1): [0, 2, 4, 5]})


def cfg3_model_kwargs():
    # BASELINE config 3: binary, order 4, maxl [12,8,2], 10 radial functions, model_type 3 (F = 9385)
    return dict(n_type=2, cutoff=6.0, model_type=3, max_p=2, gtinv_order=4, gtinv_maxl=[12, 8, 2], n_gaussians=10)


def cfg4_model_kwargs():
    # BASELINE config 4: ternary, same gtinv settings (F = 45090)
    return dict(n_type=3, cutoff=6.0, model_type=3, max_p=2, gtinv_order=4, gtinv_maxl=[12, 8, 2], n_gaussians=10)


def ternary_p3_model_kwargs():
    # small on purpose: max_p = 3 models grow as n_linear^3 / 6
    return dict(n_type=3, cutoff=4.5, model_type=2, max_p=3, gtinv_order=2, gtinv_maxl=[0],
                pair_params=[[1.0, 0.0], [0.0, 0.0]])


def pair_model_kwargs(n_type=2, model_type=2, max_p=2):
    # feature_type = "pair" (radial sums only); binary: conditional radial sets as in binary_model_kwargs
    cond = {(0, 0): [0, 1, 2, 5], (0, 1): [1, 2, 3, 4, 5], (1, 1): [0, 2, 4, 5]} if n_type == 2 else None
    return dict(n_type=n_type, cutoff=5.0, model_type=model_type, max_p=max_p, gtinv_order=0, gtinv_maxl=[],
                pair_params=[[1.0, m] for m in np.linspace(0, 4, 5)] + [[0.0, 0.0]],
                pair_params_conditional=cond, feature_type="pair")


def mgo_model_kwargs(feature_type):
    # the reference's tests/test_calc/files/mlps/polymlp.yaml.{pair,gtinv}.MgO
    kw = dict(n_type=2, cutoff=8.0, max_p=2, pair_params=[[1.0, float(m)] for m in range(8)] + [[0.0, 0.0]])
    if feature_type == "pair":
        return dict(kw, model_type=2, gtinv_order=0, gtinv_maxl=[], feature_type="pair")
    return dict(kw, model_type=3, gtinv_order=3, gtinv_maxl=[4, 4])


def load_mgo():
    return np.load(os.path.join(GOLDEN, "mgo.npz"))


def skewed_cell(n_type, n_atom=9, seed=1):
    rng = np.random.default_rng(seed)
    axis = np.array([[5.0, 2.6, 0.3], [0.1, 4.5, 2.4], [0.0, 0.2, 5.5]])
    pos = rng.random((3, n_atom))
    types = rng.integers(0, n_type, n_atom)
    types[:n_type] = np.arange(n_type)
    return axis, axis @ pos, types.astype(np.int32)


def small_skewed_cell(seed=7):
    """Small cell whose lattice dot products are < 1 in magnitude: exercises the reference's
    integer-abs branch of NeighborCell::compute_metric and many periodic images."""
    rng = np.random.default_rng(seed)
    axis = np.array([[1.3, 0.9, 0.1], [0.0, 1.1, 0.7], [0.0, 0.0, 1.6]]) * 1.7
    pos = rng.random((3, 3))
    return axis, axis @ pos, np.zeros(3, np.int32)


def fcc_supercell(rep=(4, 4, 4), a=4.05, sigma=0.05, seed=20240):
    """BASELINE config 2 structures: fcc conventional cells, Gaussian displacements."""
    base = np.array([[0, 0, 0], [0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]], float)
    cells = np.array([[i, j, k] for i in range(rep[0]) for j in range(rep[1]) for k in range(rep[2])], float)
    frac = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) / np.array(rep, float)
    axis = np.diag(np.array(rep, float) * a)
    rng = np.random.default_rng(seed)
    pc = axis @ frac.T + rng.normal(0.0, sigma, size=(3, frac.shape[0]))
    return axis, pc, np.zeros(frac.shape[0], np.int32)


def bcc_supercell(rep=(3, 3, 2), a=3.2, n_type=2, sigma=0.05, seed=5):
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]], float)
    cells = np.array([[i, j, k] for i in range(rep[0]) for j in range(rep[1]) for k in range(rep[2])], float)
    frac = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) / np.array(rep, float)
    axis = np.diag(np.array(rep, float) * a)
    rng = np.random.default_rng(seed)
    pc = axis @ frac.T + rng.normal(0.0, sigma, size=(3, frac.shape[0]))
    types = rng.integers(0, n_type, frac.shape[0]).astype(np.int32)
    types[:n_type] = np.arange(n_type)
    return axis, pc, types


def cfg4_small_cell():
    """16-atom ternary bcc cell for the config-4 model goldens (all three centre types, all six type pairs)."""
    return bcc_supercell(rep=(2, 2, 2), a=3.2, n_type=3, seed=8)


def projection_matrix(n_features, n_proj=16, seed=4):
    """Seeded Gaussian matrix for the projected-row goldens of models whose rows are too wide to store in full."""
    return np.random.default_rng(seed).normal(size=(n_features, n_proj))


def load_si_dataset():
    """The reference's bundled Si-64 phono3py training set (tests/files/phonopy_training_dataset.yaml.xz),
    stored as arrays by tests/golden/make_golden.py."""
    d = np.load(os.path.join(GOLDEN, "si_dataset.npz"))
    axis = d["axis"]
    inv = np.linalg.inv(axis)
    positions_c = [axis @ (d["frac"] + inv @ disp.T) for disp in d["displacements"]]
    return axis, positions_c, d["forces"], d["energies"]


def split_ids_train_test(n_data, train_ratio=0.9):
    """src/pypolymlp/core/utils.py:18-27 of the reference."""
    n_train = round(n_data * train_ratio)
    n_test = n_data - n_train
    test_ids = np.round(np.linspace(0, n_data - 1, num=n_test)).astype(int)
    mask = np.ones(n_data, dtype=bool)
    mask[test_ids] = False
    return np.where(mask)[0], test_ids


def x_rel_err(a, b):
    """Parity metric for design-matrix rows: |a - b| relative to the largest entry of the same column
    (near-zero entries are cancellations), with the column scale floored at 1e-8 of the global maximum
    (columns that are identically ~1e-20 carry no information; the reference itself drops products below
    1e-20, polymlp_features.cpp:186)."""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    scale = np.maximum(np.abs(b).max(axis=0), 1e-8 * max(np.abs(b).max(), 1e-300))
    return float((np.abs(a - b) / scale).max())
