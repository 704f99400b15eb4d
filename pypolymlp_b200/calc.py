"""Energy / force / stress evaluation from potential FILES, single or hybrid: the host-side composition the reference
does above the pybind11 boundary, on top of the device evaluation (`libmlpcpp.PotentialPropertiesFast`).

Reference: src/pypolymlp/calculator/properties.py:20-60 (Properties: one file -> PropertiesSingle, several ->
PropertiesHybrid), properties_single.py:16-140 (type mapping, `type_full = False` sub-models that see only the atoms of
their own elements, forces scattered back to the full atom list), properties_hybrid.py:14-62 (sum over sub-models),
calculator/utils/properties_utils.py:9-42 (find_active_atoms).

A structure is (axis 3x3 with the lattice vectors as columns, positions_c 3xN Cartesian, elements: N strings).
Returns follow the reference: energy eV/cell, forces (3, N) eV/A, stress 6 virial components (xx, yy, zz, xy, yz, zx)
in eV/cell.  There is no host evaluation here: every sub-model is a `PotentialPropertiesFast` and needs a CUDA device."""

import numpy as np

from .io_legacy import load_mlp
from .libmlpcpp import PotentialPropertiesFast


class PropertiesSingle:
    """One potential (properties_single.py:16-140)."""

    def __init__(self, pot=None, params_dict=None, coeffs=None, elements=None, type_full=True, device=None):
        if pot is not None:
            params_dict, coeffs, meta = load_mlp(pot)
            elements, type_full = meta["elements"], meta["type_full"]
        if params_dict is None or coeffs is None or elements is None:
            raise ValueError("a potential file, or params_dict + coeffs + elements, is required")
        self.elements = [str(e) for e in elements]
        if len(set(self.elements)) != len(self.elements):
            raise RuntimeError("Not available for system with spin configurations.")  # properties_utils.py:15-16
        self.type_full = bool(type_full) or type_full is None
        self.params_dict, self.coeffs = params_dict, np.asarray(coeffs, np.float64)
        self._obj = PotentialPropertiesFast(params_dict, self.coeffs, device=device)

    def _active(self, elements):
        """Atoms this model sees and their types.  A full model must know every element of the structure."""
        elements = [str(e) for e in elements]
        if self.type_full:
            unknown = sorted(set(elements) - set(self.elements))
            if unknown:
                raise ValueError("elements %s are not part of the potential (%s)" % (unknown, self.elements))
            atoms = np.arange(len(elements))
        else:
            atoms = np.array([i for i, e in enumerate(elements) if e in self.elements], dtype=int)
        types = np.array([self.elements.index(elements[i]) for i in atoms], dtype=np.int32)
        return atoms, types

    def eval(self, axis, positions_c, elements):
        e, f, s = self.eval_multiple([axis], [positions_c], [elements])
        return float(e[0]), f[0], s[0]

    def eval_multiple(self, axis_array, positions_c_array, elements_array):
        n_st = len(axis_array)
        energies, stresses = np.zeros(n_st), np.zeros((n_st, 6))
        forces = [np.zeros((3, len(el))) for el in elements_array]
        picked = [self._active(el) for el in elements_array]
        live = [k for k in range(n_st) if len(picked[k][0]) > 0]   # structures without active atoms contribute zero
        if not live:
            return energies, forces, stresses
        self._obj.eval_multiple([np.asarray(axis_array[k], float) for k in live],
                                [np.ascontiguousarray(np.asarray(positions_c_array[k], float)[:, picked[k][0]])
                                 for k in live],
                                [picked[k][1] for k in live])
        e, f, s = self._obj.get_e_array(), self._obj.get_f_array(), self._obj.get_s_array()
        for i, k in enumerate(live):
            energies[k] = e[i]
            stresses[k] = s[i]
            forces[k][:, picked[k][0]] = np.asarray(f[i]).T
        return energies, forces, stresses


class Properties:
    """Properties(pot=file) or Properties(pot=[file, ...]) for hybrid models (properties.py:20-60,
    properties_hybrid.py:14-62): the sub-models' energies, forces and stresses add up."""

    def __init__(self, pot, device=None):
        pots = list(pot) if isinstance(pot, (list, tuple)) else [pot]
        if not pots:
            raise ValueError("no potential file given")
        self._props = [PropertiesSingle(pot=p, device=device) for p in pots]

    @property
    def elements(self):
        return self._props[0].elements

    def eval(self, axis, positions_c, elements):
        e, f, s = self.eval_multiple([axis], [positions_c], [elements])
        return float(e[0]), f[0], s[0]

    def eval_multiple(self, axis_array, positions_c_array, elements_array):
        energies, forces, stresses = self._props[0].eval_multiple(axis_array, positions_c_array, elements_array)
        for prop in self._props[1:]:
            e1, f1, s1 = prop.eval_multiple(axis_array, positions_c_array, elements_array)
            energies += e1
            stresses += s1
            for k, fk in enumerate(f1):
                forces[k] += fk
        return energies, forces, stresses
