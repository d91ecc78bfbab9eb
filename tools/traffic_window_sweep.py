import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import scalesim_b200 as ssb
from scalesim_b200 import traffic
sc = traffic.synthetic_grid(1024, 1024, 10_000_000, seed=1)
ev = traffic.pack_events(sc)
per = 10_000_000
for bucket, window in ((11, 1), (11, 2), (11, 4), (22, 1), (11, 8)):
    eng = traffic.make_engine(sc, bucket_ticks=bucket, window_buckets=window, max_events=int(per*2)+(1<<20), max_batch=per+(1<<16), load=False)
    best = 1e9
    for it in range(3):
        eng.reset(); eng.load_events(ev)
        st = eng.run()
        best = min(best, st.loop_ms)
    print(bucket, window, "loop ms", round(best,2), "rounds", st.rounds, "committed", st.committed, "rollbacks", st.rollbacks, "antis", st.antis_sent, flush=True)
    eng.close()
