"""Generates tests/golden/deskew_np32.npz -- known answers for imageProjection's projectPointCloud / deskewPoint
(src/imageProjection.cpp:573-673), computed WITHOUT the oracle: python floats for the IMU integration and the rotation
look-up (IEEE double, evaluated expression by expression), numpy float32 scalars for every single-precision operation,
in the operation order SURVEY.md A.3 and DESIGN.md section 4 state (pcl::getTransformation with sin / cos in double
rounded to float; Eigen's Affine3f inverse by cofactors and Affine3f product with 3-term sums as a0 + (a1 + a2); the
point left to right).  The reference ships no fixture for this step, so this is an independent restatement of the same
specification; oracle/deskew_oracle.cpp and the device kernel must both reproduce it bit for bit.

Run from the repo root:  python tests/golden/make_golden_deskew.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def get_transformation(roll, pitch, yaw):
    A, B = f32(np.cos(np.float64(yaw))), f32(np.sin(np.float64(yaw)))
    C, D = f32(np.cos(np.float64(pitch))), f32(np.sin(np.float64(pitch)))
    E, F = f32(np.cos(np.float64(roll))), f32(np.sin(np.float64(roll)))
    DE, DF = D * E, D * F
    return np.array([[A * C, A * DF - B * E, B * F + A * DE], [B * C, A * E + B * DF, B * DE - A * F], [-D, C * F, C * E]], f32)


def sum3(a, b, c):
    return f32(a + f32(b + c))


def inverse3(m):
    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return f32(f32(m[i1, j1] * m[i2, j2]) - f32(m[i1, j2] * m[i2, j1]))
    c0 = [cof(k, 0) for k in range(3)]
    det = sum3(c0[0] * m[0, 0], c0[1] * m[1, 0], c0[2] * m[2, 0])
    inv = f32(f32(1.0) / det)
    r = np.empty((3, 3), f32)
    for k in range(3):
        r[0, k] = c0[k] * inv
        r[1, k] = cof(k, 1) * inv
        r[2, k] = cof(k, 2) * inv
    return r


def main():
    rng = np.random.default_rng(20261020)
    n = 600
    pts = (rng.normal(size=(n, 4)) * np.array([25, 25, 2.5, 40])).astype(f32)
    pts[5] = (0.3, 0.1, 0.2, 1.0)                       # inside the minimum range
    ring = rng.integers(-1, 18, n).astype(np.int32)     # some rings outside [0, 16)
    time = np.sort(rng.uniform(0.0, 0.0995, n)).astype(f32)
    time[-3:] = (0.2, 0.3, 0.4)                         # later than the last IMU sample
    t_cur = 1234.5
    stamps = [t_cur - 0.006 + 0.002 * k for k in range(60)]
    gyro = rng.normal(0, 0.5, (60, 3))
    # imuDeskewInfo (:441-496), python doubles
    imu_t, imu_r = [stamps[0]], [(0.0, 0.0, 0.0)]
    t_end = t_cur + float(time[-1])
    for k in range(1, 60):
        if stamps[k] > t_end + 0.01:
            break
        dt = stamps[k] - imu_t[-1]
        imu_r.append(tuple(imu_r[-1][a] + float(gyro[k][a]) * dt for a in range(3)))
        imu_t.append(stamps[k])
    last = len(imu_t) - 1
    params = dict(min_range=1.0, max_range=60.0, n_scan=16, downsample_rate=2, point_filter_num=3)

    def rot_at(rel):
        pt = t_cur + float(rel)
        front = 0
        while front < last:
            if pt < imu_t[front]:
                break
            front += 1
        if pt > imu_t[front] or front == 0:
            return [f32(imu_r[front][a]) for a in range(3)]
        back = front - 1
        rf = (pt - imu_t[back]) / (imu_t[front] - imu_t[back])
        rb = (imu_t[front] - pt) / (imu_t[front] - imu_t[back])
        return [f32(imu_r[front][a] * rf + imu_r[back][a] * rb) for a in range(3)]

    out, start_inv = [], None
    with np.errstate(all="ignore"):
        for i in range(n):
            x, y, z, w = pts[i]
            rng_ = f32(np.sqrt(f32(f32(f32(x * x) + f32(y * y)) + f32(z * z))))
            if rng_ < f32(params["min_range"]) or rng_ > f32(params["max_range"]):
                continue
            if ring[i] < 0 or ring[i] >= params["n_scan"] or ring[i] % params["downsample_rate"] != 0 or i % params["point_filter_num"] != 0:
                continue
            Tf = get_transformation(*rot_at(time[i]))
            if start_inv is None:
                start_inv = inverse3(Tf)
                t_inv = [sum3(f32(-start_inv[a, 0]) * f32(0), f32(-start_inv[a, 1]) * f32(0), f32(-start_inv[a, 2]) * f32(0)) for a in range(3)]
            bt = np.empty((3, 3), f32)
            for a in range(3):
                for b in range(3):
                    bt[a, b] = sum3(start_inv[a, 0] * Tf[0, b], start_inv[a, 1] * Tf[1, b], start_inv[a, 2] * Tf[2, b])
            tt = [f32(sum3(start_inv[a, 0] * f32(0), start_inv[a, 1] * f32(0), start_inv[a, 2] * f32(0)) + t_inv[a]) for a in range(3)]
            o = [f32(f32(f32(f32(bt[a, 0] * x) + f32(bt[a, 1] * y)) + f32(bt[a, 2] * z)) + tt[a]) for a in range(3)]
            out.append([o[0], o[1], o[2], w])
    out = np.array(out, f32)
    np.savez_compressed(os.path.join(HERE, "deskew_np32.npz"), pts=pts, ring=ring, time=time, imu_time=np.array(imu_t), imu_rot=np.array(imu_r),
                        time_scan_cur=t_cur, out=out, **{k: np.array(v) for k, v in params.items()})
    print("deskew_np32.npz:", len(out), "of", n, "points survive; IMU samples", len(imu_t))


if __name__ == "__main__":
    main()
