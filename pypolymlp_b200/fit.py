"""Host side of the fit path above the C ABI: weights, batching, scales, ridge solve.

Mirrors the reference's Python driver for this path with the same names and argument meaning:
  apply_weights            src/pypolymlp/mlp_dev/core/utils_weights.py:10-98
  calc_xtx_xty             src/pypolymlp/mlp_dev/core/data_sequential.py:25-156
  compute_scales           src/pypolymlp/mlp_dev/core/utils_scales.py:6-40
  solver_ridge             src/pypolymlp/mlp_dev/standard/solvers.py:9-84
  compute_rmse (from XtX)  src/pypolymlp/mlp_dev/core/utils_model_selection.py:37-72
The O(rows) / O(F^2) host arithmetic stays on the host as in the reference; the feature build and
X^T X accumulation (the hot path) run on the GPU through PotentialXtX.  There is no CPU fallback.
"""

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .libmlpcpp import PotentialXtX, StructureBatch, comm_unique_id


@dataclass
class Dataset:
    """The fields of pypolymlp.core.dataset.Dataset that the fit path reads."""

    axis: list
    positions_c: list            # (3, N) Cartesian per structure
    types: list
    energies: np.ndarray         # (n_st,)
    forces: Optional[np.ndarray] = None    # concatenated, atom-major x,y,z per structure
    stresses: Optional[np.ndarray] = None  # concatenated, 6 per structure (xx,yy,zz,xy,yz,zx)
    include_force: bool = True
    include_stress: bool = False
    weight: float = 1.0
    name: str = "dataset"
    total_n_atoms: np.ndarray = field(default=None)

    def __post_init__(self):
        self.energies = np.asarray(self.energies, dtype=float)
        if self.total_n_atoms is None:
            self.total_n_atoms = np.array([np.asarray(p).shape[1] for p in self.positions_c])
        n6 = 6 * len(self.energies)
        if self.include_force:
            if self.forces is None:
                raise ValueError("forces are required when include_force is True")
            self.forces = np.asarray(self.forces, dtype=float).reshape(-1)
            if self.stresses is None:
                self.stresses = np.zeros(n6)
            self.stresses = np.asarray(self.stresses, dtype=float).reshape(-1)

    def slice(self, begin, end):
        fo = np.concatenate([[0], np.cumsum(3 * self.total_n_atoms)])
        return Dataset(
            self.axis[begin:end], self.positions_c[begin:end], self.types[begin:end], self.energies[begin:end],
            None if not self.include_force else self.forces[fo[begin]:fo[end]],
            None if not self.include_force else self.stresses[6 * begin:6 * end],
            self.include_force, self.include_stress, self.weight, self.name, self.total_n_atoms[begin:end])


def _set_weight_energy_data(energy, total_n_atoms, min_e=None):
    e_per_atom = energy / total_n_atoms
    if min_e is None:
        min_e = np.min(e_per_atom)
    weight_e = np.ones(len(energy))
    weight_e[e_per_atom > min_e * 0.75] = 0.5
    weight_e[e_per_atom > min_e * 0.50] = 0.3
    weight_e[e_per_atom > 0.0] = 0.1
    return weight_e


def _set_weight_force_data(forces, tol=1e-12):
    weight_f = np.abs(forces)
    weight_f[weight_f < tol] = tol
    weight_f = np.reciprocal(weight_f)
    weight_f[weight_f > 1.0] = 1.0
    return weight_f


def _set_weight_stress_data(stress, weight_stress, tol=1e-12):
    nonzero = np.abs(stress) > tol
    log1 = np.ones(len(stress)) * np.log10(tol)
    log1[nonzero] = np.log10(np.abs(stress)[nonzero])
    weight_s = np.power(5, -log1)
    weight_s[weight_s > 1.0] = 1.0
    return weight_s * weight_stress


def apply_weights(dataset: Dataset, weight_stress=0.1, min_e=None):
    """Row weights w and weighted targets y of one (sliced) dataset in the PyModel row layout
    [energies | stress | forces]; the X rows are scaled by w on the GPU."""
    n_st = len(dataset.energies)
    we = _set_weight_energy_data(dataset.energies, dataset.total_n_atoms, min_e=min_e) * dataset.weight
    w, y = [we], [we * dataset.energies]
    if dataset.include_force:
        if dataset.include_stress:
            ws = _set_weight_stress_data(dataset.stresses, weight_stress * dataset.weight)
            w.append(ws)
            y.append(ws * dataset.stresses)
        else:  # rows exist but are zeroed (utils_weights.py:94-97)
            w.append(np.zeros(6 * n_st))
            y.append(np.zeros(6 * n_st))
        wf = _set_weight_force_data(dataset.forces) * dataset.weight
        w.append(wf)
        y.append(wf * dataset.forces)
    return np.concatenate(w), np.concatenate(y)


def get_min_energy(datasets):
    """src/pypolymlp/mlp_dev/core/utils.py:10-22."""
    min_e = 1e10
    for data in datasets:
        if len(data.energies) == 0:
            raise RuntimeError("Empty energy data.")
        min_e = min(min_e, np.min(data.energies / data.total_n_atoms))
    return min_e


def get_batch_slice(n_data, batch_size):
    begin = list(range(0, n_data, batch_size))
    end = list(begin[1:]) + [n_data] if len(begin) > 1 else [n_data]
    return begin, end


def compute_scales(scales, xe_sum, xe_sq_sum, n_data, include_force=True, threshold=1e-10):
    if scales is None:
        variance = xe_sq_sum / n_data - np.square(xe_sum / n_data)
        variance[variance < 0.0] = 1.0
        scales = np.sqrt(variance)
    zero_ids = np.abs(scales) < (threshold if include_force else threshold * threshold)
    scales = scales.copy()
    scales[zero_ids] = 1.0
    return scales, zero_ids


@dataclass
class PolymlpDataXY:
    xtx: Optional[np.ndarray] = None
    xty: Optional[np.ndarray] = None
    scales: Optional[np.ndarray] = None
    xe_sum: Optional[np.ndarray] = None
    xe_sq_sum: Optional[np.ndarray] = None
    y_sq_norm: float = 0.0
    total_n_data: int = 0
    min_energy: Optional[float] = None


def accumulate_datasets(acc: PotentialXtX, datasets, min_energy, weight_stress=0.1, batch_size=64):
    """Feeds every dataset to the GPU accumulator in structure batches."""
    for data in datasets:
        n_str = len(data.energies)
        for begin, end in zip(*get_batch_slice(n_str, batch_size)):
            sl = data.slice(begin, end)
            w, y = apply_weights(sl, weight_stress=weight_stress, min_e=min_energy)
            acc.add(sl.axis, sl.positions_c, sl.types, [sl.include_force] * (end - begin), w, y)


def finalize_xtx_xty(res: dict, datasets, scales=None, min_energy=None, scale_threshold=1e-10):
    """Tail of calc_xtx_xty (data_sequential.py:72-92): scales, zeroing, normalisation."""
    n_data = sum(len(d.energies) for d in datasets)
    include_force = any(d.include_force for d in datasets)
    scales, zero_ids = compute_scales(scales, res["xe_sum"], res["xe_sq_sum"], n_data,
                                      include_force=include_force, threshold=scale_threshold)
    xtx, xty = res["xtx"], res["xty"].copy()
    xtx[zero_ids] = 0.0
    xtx[:, zero_ids] = 0.0
    xty[zero_ids] = 0.0
    xtx /= scales[:, np.newaxis]
    xtx /= scales[np.newaxis, :]
    xty /= scales
    return PolymlpDataXY(xtx=xtx, xty=xty, scales=scales, xe_sum=res["xe_sum"], xe_sq_sum=res["xe_sq_sum"],
                         y_sq_norm=res["y_sq_norm"], total_n_data=res["total_n_data"], min_energy=min_energy)


def calc_xtx_xty(params_dict, datasets, scales=None, min_energy=None, weight_stress=0.1, batch_size=64,
                 scale_threshold=1e-10, device=None, flags=0):
    """Compute X.T @ X and X.T @ y on the GPU (same outputs as the reference's calc_xtx_xty)."""
    if min_energy is None:
        min_energy = get_min_energy(datasets)
    acc = PotentialXtX(params_dict, device=device, flags=flags)
    accumulate_datasets(acc, datasets, min_energy, weight_stress, batch_size)
    return finalize_xtx_xty(acc.finalize(), datasets, scales, min_energy, scale_threshold)


def solver_ridge(xtx, xty, alphas=(1e-3, 1e-2, 1e-1)):
    """Ridge regression by Cholesky (posv) per alpha with incremental diagonal update."""
    from scipy.linalg.lapack import get_lapack_funcs

    (posv,) = get_lapack_funcs(("posv",), (xtx, xty))
    n = xtx.shape[0]
    coefs = np.zeros((n, len(alphas)))
    prev = 0.0
    for i, alpha in enumerate(alphas):
        xtx.flat[:: n + 1] += alpha - prev
        _, x, info = posv(xtx.T if xtx.flags["C_CONTIGUOUS"] else xtx, xty, lower=False,
                          overwrite_a=False, overwrite_b=False)
        coefs[:, i] = x if info == 0 else np.ones(n) * 1e30
        prev = alpha
    xtx.flat[:: n + 1] -= prev
    return coefs


def compute_rmse(coefs_array, data_xy: PolymlpDataXY):
    """RMSE from X.T X: c^T (XtX) c - 2 c^T Xty + y^T y (utils_model_selection.py:37-72)."""
    out = []
    for c in coefs_array.T:
        mse = (c @ (data_xy.xtx @ c) - 2 * c @ data_xy.xty + data_xy.y_sq_norm) / data_xy.total_n_data
        out.append(np.sqrt(mse) if mse >= 0 else 1e10)
    rm = np.array(out)
    if np.all(rm > 1e6):
        raise RuntimeError("Matrix (X.T @ X + alpha * I) may be singular. "
                           "This singularity issue might be reduced by increasing "
                           "the value of alpha (the magnitude of the penalty term).")
    return rm


def fit(params_dict, train, test, alphas, weight_stress=0.1, batch_size=64, device=None):
    """Standard fit (src/pypolymlp/mlp_dev/standard/fit.py:12-63): returns the best model."""
    train_xy = calc_xtx_xty(params_dict, train, weight_stress=weight_stress, batch_size=batch_size, device=device)
    coefs = solver_ridge(train_xy.xtx, train_xy.xty, alphas=alphas)
    rmse_train = compute_rmse(coefs, train_xy)
    test_xy = calc_xtx_xty(params_dict, test, scales=train_xy.scales, min_energy=train_xy.min_energy,
                           weight_stress=weight_stress, batch_size=batch_size, device=device)
    rmse_test = compute_rmse(coefs, test_xy)
    idx = int(np.argmin(rmse_test))
    return {"coeffs": coefs[:, idx], "scales": train_xy.scales, "alpha": alphas[idx],
            "rmse_train": rmse_train[idx], "rmse_test": rmse_test[idx], "coefs_array": coefs,
            "rmse_train_array": rmse_train, "rmse_test_array": rmse_test}


def fit_device(params_dict, train, test, alphas, weight_stress=0.1, batch_size=64, device=None):
    """Same as fit(), with the ridge solve of the training set on the GPU (cuSOLVER Cholesky on the
    device-resident accumulator; pm_fit_solve_ridge).  The O(F^2) test-set RMSE stays on the host."""
    min_energy = get_min_energy(train)
    acc = PotentialXtX(params_dict, device=device)
    accumulate_datasets(acc, train, min_energy, weight_stress, batch_size)
    n_energy = sum(len(d.energies) for d in train)
    include_force = any(d.include_force for d in train)
    scales, coefs, rmse_train = acc.solve_ridge(alphas, n_energy, include_force=include_force)
    test_xy = calc_xtx_xty(params_dict, test, scales=scales, min_energy=min_energy, weight_stress=weight_stress,
                           batch_size=batch_size, device=device)
    rmse_test = compute_rmse(coefs, test_xy)
    idx = int(np.argmin(rmse_test))
    return {"coeffs": coefs[:, idx], "scales": scales, "alpha": alphas[idx], "rmse_train": rmse_train[idx],
            "rmse_test": rmse_test[idx], "coefs_array": coefs, "rmse_train_array": rmse_train,
            "rmse_test_array": rmse_test}


# ---- multi-GPU: structures shard across ranks, one NCCL reduce of the packed accumulator ----------
def compute_error(params_dict, scaled_coeffs, dataset: Dataset, stress_unit="eV", device=None, prop=None):
    """Prediction errors of a fitted model on one dataset, through the device eval path with contiguous buffers.

    Mirrors PolymlpEvalAccuracy.compute_error_single (PY/mlp_dev/core/eval_accuracy.py:49-140): RMSE / MAE of the
    energy per atom, of the force components and of the stress (per atom in eV, or in GPa from the cell volumes).
    scaled_coeffs = coefs / scales, as written to polymlp.yaml.  Returns the reference's error_dict keys."""
    from .libmlpcpp import PotentialPropertiesFast

    if prop is None:
        prop = PotentialPropertiesFast(params_dict, scaled_coeffs, device=device)
    e, f_list, s = prop._run(dataset.axis, dataset.positions_c, dataset.types)
    n_atoms = np.asarray(dataset.total_n_atoms, dtype=float)

    def _err(true, pred, normalize=None):
        true, pred = np.asarray(true, dtype=float).reshape(-1), np.asarray(pred, dtype=float).reshape(-1)
        if normalize is not None:
            true, pred = true / normalize, pred / normalize
        d = true - pred
        return float(np.sqrt(np.mean(np.square(d)))), float(np.mean(np.abs(d)))

    rmse_e, mae_e = _err(dataset.energies, e, n_atoms)
    out = {"energy": rmse_e, "energy_mae": mae_e, "force": None, "force_mae": None, "stress": None, "stress_mae": None}
    if dataset.include_force and dataset.forces is not None:
        pred_f = np.concatenate([np.asarray(f).reshape(-1) for f in f_list]) if f_list else np.zeros(0)
        out["force"], out["force_mae"] = _err(dataset.forces, pred_f)
    if dataset.include_stress and dataset.stresses is not None:
        if stress_unit == "GPa":
            vol = np.array([abs(np.linalg.det(np.asarray(a, dtype=float))) for a in dataset.axis])
            norm = np.repeat(vol, 6) / 160.21766208
        else:
            norm = np.repeat(n_atoms, 6)
        out["stress"], out["stress_mae"] = _err(dataset.stresses, s, norm)
    return out


def reduce_accumulator(acc: PotentialXtX, dst=0):
    """Sum the per-rank partial [C upper tiles | xe_sum | xe_sq_sum | n_data] onto rank `dst`: one ncclReduce issued by
    the library itself (pm_fit_reduce); needs acc.comm_init_rank / comm_init_from_env first (no-op for one rank)."""
    acc.reduce(dst)


def comm_init_from_env(acc: PotentialXtX, timeout_s=120.0):
    """One process per GPU under torchrun / mpirun-style launchers (RANK, WORLD_SIZE, MASTER_PORT in the environment):
    rank 0 creates the NCCL unique id and publishes it through a file in the node's temp directory, the others pick it
    up.  Single node only (which is what the C ABI's pm_multi / the bench cover); returns (rank, world_size)."""
    import os
    import tempfile
    import time

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    tag = "%s_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x"), os.getppid())
    path = os.path.join(tempfile.gettempdir(), f"polymlp_b200_ncclid_{tag}.bin")
    if rank == 0:
        uid = comm_unique_id()
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
    else:
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > timeout_s:
                raise TimeoutError(f"rank {rank}: no NCCL id at {path}")
            time.sleep(0.01)
        with open(path, "rb") as f:
            uid = f.read()
    acc.comm_init_rank(world, rank, uid)
    acc.barrier()
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return rank, world


def shard_range(n, rank, world_size):
    """Contiguous, balanced slice of n structures for `rank`."""
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
