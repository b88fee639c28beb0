"""tools/heat_deep_probe.py -- deep 2-D tiles vs row slabs vs one field, all on ONE GPU (several ranks in one process): where
does the time of a deep-tile launch go? python tools/heat_deep_probe.py [launches]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alpaka_b200 as ab
from alpaka_b200 import multi

launches = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = ab.Platform().get_dev_by_idx(0)
NY, NX, G = 8192, 8192, 4
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)


def run(name, make, world):
    queues = [ab.Queue(dev) for _ in range(world)]
    rs = [make(q, r) for r, q in enumerate(queues)]
    multi.connect_in_process(rs)
    for r in rs:
        r.upload(r.initial_field())
    def go(n):
        for _ in range(n):
            for r in rs:
                r.step(G)
    go(3)
    for q in queues: q.wait()
    ab.enqueue(queues[0], e0)
    go(launches)
    for q in queues[1:]: q.wait()
    ab.enqueue(queues[0], e1); queues[0].wait()
    for q in queues: q.wait()
    print(f"{name:42s}: {e0.elapsed_ms(e1) / launches * 1e3:8.1f} us per launch of {G} levels (all ranks on one GPU)", flush=True)
    for r in rs:
        assert r.status() == 0
        r.close()


run("one field 8192x8192 (slab, world 1)", lambda q, r: multi.HeatSlab(q, r, 1, NY, NX, levels=G), 1)
run("2 row slabs of 4096x8192", lambda q, r: multi.HeatSlab(q, r, 2, NY, NX, levels=G), 2)
run("deep tiles 1x1", lambda q, r: multi.HeatTileDeep(q, r, 1, NY, NX, levels=G, grid=(1, 1)), 1)
run("deep tiles 2x1 (rows only)", lambda q, r: multi.HeatTileDeep(q, r, 2, NY, NX, levels=G, grid=(2, 1)), 2)
run("deep tiles 1x2 (columns only)", lambda q, r: multi.HeatTileDeep(q, r, 2, NY, NX, levels=G, grid=(1, 2)), 2)
run("deep tiles 2x2", lambda q, r: multi.HeatTileDeep(q, r, 4, NY, NX, levels=G, grid=(2, 2)), 4)
