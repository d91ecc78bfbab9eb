"""Host-side scenario ingestion timing (profiles/r2_ingest_host.json): python tools/ingest_bench.py [n_trips]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scalesim_b200 import traffic, scenario
n_trips = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
import tempfile
d = tempfile.mkdtemp(prefix='ssb_ingest_')
t = time.time(); sc = traffic.synthetic_grid(1024, 1024, n_trips, seed=1); sc.partition = traffic.stripe_partition(1024, 1024, 8); gen = time.time() - t
rd, trip, part, img = (os.path.join(d, n) for n in ('rd.csv', 'trip.csv', 'part', 'grid.ssbscn'))
t = time.time(); scenario.write_traffic_csv(sc, rd, trip, part); wr = time.time() - t
size = sum(os.path.getsize(p) for p in (rd, trip, part))
res = {'trips': n_trips, 'csv_bytes': size, 'generate_s': round(gen, 2), 'write_csv_s': round(wr, 2)}
for th in (1, 4, 0):
    t = time.time(); nat = scenario.read_traffic_csv(rd, trip, part, threads=th); res[f'native_read_s_threads_{th or os.cpu_count()}'] = round(time.time() - t, 3)
assert np.array_equal(nat['route_nodes'], sc.route_nodes) and np.array_equal(nat['trip_dep'], sc.trip_dep) and np.array_equal(nat['road_len'], sc.road_len)
t = time.time(); scenario.compile_traffic_csv(rd, trip, img, partition_file=part); res['compile_to_image_s'] = round(time.time() - t, 3)
res['image_bytes'] = os.path.getsize(img)
t = time.time()
with scenario.Image.open(img) as im:
    n = im.n_events; s = int(im.events['p1'][:: 4096].sum())
res['open_image_s'] = round(time.time() - t, 4)
if n_trips <= 200_000:
    t = time.time(); traffic.read_scenario(rd, trip, part); res['python_line_reader_s'] = round(time.time() - t, 2)
print(json.dumps(res))
