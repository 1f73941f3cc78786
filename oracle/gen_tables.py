# Runs the REAL reference utils/so3.py and utils/torus.py (pure numpy/scipy) to produce the score-norm tables.
import sys, time, numpy as np
sys.path.insert(0, '/root/reference')
t0 = time.time()
np.random.seed(0)
import utils.torus as torus
print('torus done', time.time() - t0, flush=True)
np.save('torus_score_norm_seed0.npy', torus.score_norm_)
import utils.so3 as so3
print('so3 done', time.time() - t0, flush=True)
np.save('so3_exp_score_norms.npy', so3._exp_score_norms)
