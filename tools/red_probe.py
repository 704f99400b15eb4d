import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import cases
from pypolymlp_b200.libmlpcpp import PotentialXtX
from pypolymlp_b200.params import make_params_dict
acc = PotentialXtX(make_params_dict(**cases.si_model_kwargs()))
for n in (768, 768*46, 768*256):
    print("rows", n, "RED.F64 G/s:", acc.context.microbench(4, n))
