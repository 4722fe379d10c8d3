"""Where the host time of an Operator_1DtMOR epoch goes between two vn_train_batches calls: wall time of every engine call and of
the Session's own bookkeeping (run after scripts/prof_cfg3.py's set-up)."""
import sys, time, collections
import numpy as np
sys.path.insert(0, ".")
sys.argv = [sys.argv[0]]
src = open("scripts/prof_cfg3.py").read().split("epoch(); epoch()")[0]
exec(src)
from varnet_b200 import backend as be
acc = collections.defaultdict(lambda: [0.0, 0])
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            r = acc[label or name]; r[0] += time.perf_counter() - t0; r[1] += 1
    setattr(obj, name, g)
eng = next(t for t in tf.compTowers if t.local).engine
for n in ("select_table", "set_extra_inputs", "upload_bic", "train_batches_begin", "train_batches_end", "set_batch", "set_weights", "generate_table", "upload_table"):
    if hasattr(eng, n):
        wrap(eng, n)
wrap(tf.sess, "_sync_feeds"); wrap(tf.sess, "run_batches"); wrap(tf.sess, "_collect")
wrap(vn, "trainData")
epoch(); epoch()
acc.clear()
t0 = time.perf_counter()
for _ in range(20):
    epoch()
dt = time.perf_counter() - t0
print("epoch %.3f ms" % (dt / 20 * 1e3))
for k, (t, n) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-22s %8.1f us per MOR batch  (%d calls)" % (k, t / 120 * 1e6, n))
