// liloc_b200/synth/synth.cpp -- seeded synthetic LiDAR scans for tests and bench.py.
//
// The reference has no datasets, fixtures or simulators, and there is no network,
// so every input of this repo comes from here: analytic ray casting of a spinning (VLP-16 / HDL-32E)
// or non-repetitive (Mid-360-like) beam pattern against a world of axis-aligned boxes.
// Output points are in the sensor (body) frame as {x,y,z,intensity} float4, i.e. what
// laserCloudInfoHandler receives in cloud_deskewed (src/mapOptmization.cpp:253).
// Deterministic: every random number is a pure function of (seed, ray index).
//
// Build: g++ -O3 -fopenmp -shared -fPIC  (see liloc_b200/_build.py)

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double u01(uint64_t seed, uint64_t ctr, uint64_t lane) {
    const uint64_t h = mix64(mix64(seed * 0xD1342543DE82EF95ull + ctr) + lane * 0x2545F4914F6CDD1Dull);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
inline double gauss(uint64_t seed, uint64_t ctr) {
    const double a = u01(seed, ctr, 1), b = u01(seed, ctr, 2);
    return std::sqrt(-2.0 * std::log(a)) * std::cos(6.283185307179586 * b);
}

struct Rot { double m[9]; };
Rot rot_zyx(double roll, double pitch, double yaw) {
    const double A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch);
    const double E = std::cos(roll), F = std::sin(roll);
    Rot r;
    r.m[0] = A * C; r.m[1] = A * D * F - B * E; r.m[2] = B * F + A * D * E;
    r.m[3] = B * C; r.m[4] = A * E + B * D * F; r.m[5] = B * D * E - A * F;
    r.m[6] = -D;    r.m[7] = C * F;             r.m[8] = C * E;
    return r;
}

// nearest hit of ray o + t d, t > 0: boxes[0] is the room (hit from inside = exit point), the rest
// are solid obstacles (hit = entry point).  Returns t or -1.
double cast(const float* boxes, int nb, const double o[3], const double d[3]) {
    double best = -1.0;
    for (int b = 0; b < nb; ++b) {
        const float* bx = boxes + 6 * b;
        double tn = -1e300, tf = 1e300;
        bool miss = false;
        for (int a = 0; a < 3; ++a) {
            if (std::fabs(d[a]) < 1e-12) {
                if (o[a] < bx[a] || o[a] > bx[a + 3]) { miss = true; break; }
                continue;
            }
            double t1 = (bx[a] - o[a]) / d[a], t2 = (bx[a + 3] - o[a]) / d[a];
            if (t1 > t2) { const double s = t1; t1 = t2; t2 = s; }
            if (t1 > tn) tn = t1;
            if (t2 < tf) tf = t2;
        }
        if (miss || tn > tf || tf <= 0.0) continue;
        double t;
        if (b == 0) t = tf;                 // inside the room: exit through a wall / floor / ceiling
        else { if (tn <= 0.0) continue; t = tn; }
        if (t > 0.0 && (best < 0.0 || t < best)) best = t;
    }
    return best;
}

}  // namespace

extern "C" {

struct synth_sensor {
    int   pattern;      // 0 spinning, 1 random (Mid-360-like)
    int   n_rings;      // spinning: rings
    int   n_az;         // spinning: azimuth steps;  random: rays per frame
    float el_min_deg, el_max_deg;
    float range_min, range_max;
    float noise_sigma;  // range noise (m)
};

// pose6 = [roll, pitch, yaw, x, y, z] of the sensor in the world.  out must hold n_rings*n_az
// (spinning) or n_az (random) points.  Returns the number of valid returns written.
int synth_scan(const float* boxes, int n_boxes, const float* pose6, const synth_sensor* s, uint64_t seed, float* out4) {
    const Rot R = rot_zyx(pose6[0], pose6[1], pose6[2]);
    const double o[3] = {pose6[3], pose6[4], pose6[5]};
    const int total = s->pattern == 0 ? s->n_rings * s->n_az : s->n_az;
    std::vector<float> tmp((size_t)total * 4);
    std::vector<unsigned char> ok((size_t)total, 0);
    const double deg = 3.14159265358979323846 / 180.0;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < total; ++i) {
        double az, el;
        if (s->pattern == 0) {
            const int ia = i / s->n_rings, ir = i % s->n_rings;
            az = 6.283185307179586 * (double)ia / (double)s->n_az;
            el = (s->n_rings > 1) ? (s->el_min_deg + (s->el_max_deg - s->el_min_deg) * (double)ir / (double)(s->n_rings - 1)) * deg
                                  : s->el_min_deg * deg;
        } else {
            az = 6.283185307179586 * u01(seed, (uint64_t)i, 3);
            el = (s->el_min_deg + (s->el_max_deg - s->el_min_deg) * u01(seed, (uint64_t)i, 4)) * deg;
        }
        const double db[3] = {std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
        const double dw[3] = {R.m[0] * db[0] + R.m[1] * db[1] + R.m[2] * db[2],
                              R.m[3] * db[0] + R.m[4] * db[1] + R.m[5] * db[2],
                              R.m[6] * db[0] + R.m[7] * db[1] + R.m[8] * db[2]};
        double t = cast(boxes, n_boxes, o, dw);
        if (t < 0.0) continue;
        t += (double)s->noise_sigma * gauss(seed, (uint64_t)i);
        if (t < s->range_min || t > s->range_max) continue;
        tmp[4 * (size_t)i + 0] = (float)(db[0] * t);
        tmp[4 * (size_t)i + 1] = (float)(db[1] * t);
        tmp[4 * (size_t)i + 2] = (float)(db[2] * t);
        tmp[4 * (size_t)i + 3] = (float)(i & 255);
        ok[i] = 1;
    }
    int n = 0;
    for (int i = 0; i < total; ++i) if (ok[i]) { std::memcpy(out4 + 4 * (size_t)n, &tmp[4 * (size_t)i], 16); ++n; }
    return n;
}

}  // extern "C"
