import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
from dqmc_oracle import SdwOracle, SdwParams
from detqmc_b200 import DetSDWBatch
p = SdwParams(L=4, m=20, s=10, r=-1.0)
o = SdwOracle(p); b = DetSDWBatch(p, n_replicas=1)
for i in range(10):
    o.sweep_thermalization(); b.sweepThermalization()
print("after therm", np.abs(b.phi()-o.phi).max(), b.control_data().lastAccRatioLocal_phi, o.last_acc_ratio)
for i in range(10):
    o.sweep(); b.sweep()
    print(i, np.linalg.norm(o.phi[1:].mean(axis=(0,2))), np.linalg.norm(b.phi()[1:].mean(axis=(0,2))), np.abs(b.phi()-o.phi).max(), b.control_data().acceptedGlobalShifts, o.accepted_global_shifts)
