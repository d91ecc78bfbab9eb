import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import scalesim_b200 as ssb
from scalesim_b200 import traffic
sc = traffic.synthetic_grid(1024, 1024, 10_000_000, seed=1)
per = 10_000_000
eng = traffic.make_engine(sc, bucket_ticks=11, window_buckets=1, max_events=int(per*2)+(1<<20), max_batch=per+(1<<16), load=False)
ev = traffic.pack_events(sc)
for it in range(3):
    eng.reset(); eng.load_events(ev)
    if it == 2: eng.set_flags(ssb.FLAG_KERNEL_TIMING)
    st = eng.run()
    print("loop ms", round(st.loop_ms,2), "rounds", st.rounds, "committed", st.committed)
kt = eng.kernel_times()
tot = sum(v[0] for v in kt.values())
for k,(ms,n) in kt.items():
    if n: print(f"{k:16s} {ms:8.2f} ms {n:6d} launches {1000*ms/n:7.1f} us/launch {ms/tot:6.3f}")
