import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, cases
from pypolymlp_b200 import fit
from pypolymlp_b200.params import make_params_dict
from pypolymlp_b200.libmlpcpp import PotentialXtX, PotentialModel
from test_gpu_parity import _si_datasets
pd = make_params_dict(**cases.si_model_kwargs())
train_ids, test_ids = cases.split_ids_train_test(200, 0.9)
train, test = _si_datasets(train_ids[:60]), _si_datasets(test_ids)
alphas = [10.0 ** a for a in np.linspace(-3, 1, 5)]
host = fit.fit(pd, [train], [test], alphas)
dev = fit.fit_device(pd, [train], [test], alphas)
print("scales", np.abs(dev["scales"] - host["scales"]).max() / np.abs(host["scales"]).max())
print("rmse_train", np.abs(dev["rmse_train_array"] - host["rmse_train_array"]).max() / host["rmse_train_array"].max())
acc = PotentialXtX(pd)
fit.accumulate_datasets(acc, [train], fit.get_min_energy([train]))
_, coefs_dev, rmse_dev = acc.solve_ridge(alphas, len(train.energies), scales=host["scales"])
print("rmse same scales", np.abs(rmse_dev - host["rmse_train_array"]).max() / host["rmse_train_array"].max())
x = PotentialModel(pd, test.axis, test.positions_c, test.types, [20], [True], [64] * 20).get_x()
for k in range(len(alphas)):
    ph = x @ (host["coefs_array"][:, k] / host["scales"]); pdv = x @ (coefs_dev[:, k] / host["scales"])
    print(alphas[k], np.sqrt(np.mean(np.square(ph[:20] - pdv[:20]))) / 64, np.sqrt(np.mean(np.square(ph[140:] - pdv[140:]))))
