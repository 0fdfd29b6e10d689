"""Small driver used under ncu: a handful of fused steps on the C3 workload (no timing claims)."""
import sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import easyhybrid_b200 as eh
from bench import synth, make_model, B

import os
n = 1 << int(os.environ.get('EH_PROF_LOG2N', '22'))
model = make_model(eh)
xf, y = synth(n, 42)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
sess = eh.FusedSession(model, opt=eh.Adam(0.01), flags=flags)
sess.upload(0, xf, y)
sess.set_params(model.initialparameters(np.random.default_rng(0)))
sess.set_perm(np.random.default_rng(7).permutation(n))
print(sess.run_steps(B, 0, steps)[-3:])
sess.close()
