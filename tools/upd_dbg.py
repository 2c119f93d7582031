import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
from bench import WORKLOAD, ladder_values
from detqmc_b200 import DetSDWBatch
R = 64
b = DetSDWBatch(dict(WORKLOAD), n_replicas=R, rng_indices=[i + 1 for i in range(R)], r_values=ladder_values(R))
os.environ["DQMC_UPD_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "1"
acc = b.update_in_slice(100, True)
b.synchronize()
print(acc[:8])
