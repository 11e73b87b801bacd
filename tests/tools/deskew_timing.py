"""Device time of the de-skew launch (k_deskew): block 0's %globaltimer marks (start -> after the grid barrier -> end).  Usage: python tests/tools/deskew_timing.py [vlp16|hdl32|mid360]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from liloc_b200 import Context, synth  # noqa: E402


def main():
    name = (sys.argv[1] if len(sys.argv) > 1 else "hdl32").upper()
    sensor = getattr(synth, name)
    world = synth.make_world(seed=1)
    sc = synth.scan(world, np.array([0, 0, 0.3, -5, 2, 1.8], np.float32), sensor, seed=1)
    sw = synth.raw_sweep(sc, sensor, seed=1)
    t, r = synth.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    with Context(device=0, max_scan_points=1 << 18) as ctx:
        ctx.set_timing(True)
        for pf in (1, 3):
            ev, a, b = [], [], []
            for it in range(40):
                n = ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], n_scan=sensor.n_rings, point_filter_num=pf)
                mk = ctx.last_deskew_marks().astype(np.int64)
                if it >= 10:
                    ev.append(ctx.last_deskew_ms() * 1e3); a.append((mk[1] - mk[0]) * 1e-3); b.append((mk[2] - mk[1]) * 1e-3)
            print(f"{name} n_raw={len(sw['pts'])} point_filter_num={pf} n_out={n} imu={len(t)}: in-kernel total p50 {np.median(ev):.1f} us | "
                  f"marks: counts+barrier {np.median(a):.1f} us, offsets+transform {np.median(b):.1f} us")


if __name__ == "__main__":
    main()
