import numpy as np, sys
sys.path.insert(0,'/root/repo')
import latticeum_b200 as LB
from oracle import c_oracle as CO
for kappa in (8, 32):
    n=300
    A = CO.fill_uniform((kappa, n, 24), 5)
    s = LB.AjtaiCommitmentScheme.new(A)
    for count in (1,2):
        fs = CO.fill_uniform((count, n, 24), 60 + count)
        try:
            cms = s.commit_ntt_batch(fs)
            ok = all(np.array_equal(cms[k].as_ref(), CO.commit(A, fs[k])) for k in range(count))
            print(kappa, count, "ok" if ok else "MISMATCH")
        except Exception as e:
            print(kappa, count, "ERR", str(e)[:150])
    s.close()
