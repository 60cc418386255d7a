/* Synthetic-scene ray caster used by ct_icp_b200/synthetic.py (test / bench data generation only).
 * Scene: ground plane z=0, axis-aligned boxes (solid or porous "foliage"), vertical cylinders.
 * For porous boxes the return is placed at entry + an exponential free path (hash-seeded per ray),
 * dropped if it exits the box first — gives volumetric vegetation-like returns.
 * Build: gcc -O2 -fPIC -shared [-fopenmp] raycast.c -o libraycast.so -lm */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

/* boxes: nb x 7 (x0,y0,z0,x1,y1,z1, mean_free_path [0 = solid]); cyl: nc x 4 (cx,cy,r,h)
 * o,d: n x 3; out: n ranges (INFINITY when nothing is hit); boxes must be sorted by x0 (for pruning) */
void raycast_scene(const double *o, const double *d, size_t n, const double *boxes, size_t nb, const double *cyl,
                   size_t nc, double max_range, uint64_t seed, double *out) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < (long) n; ++i) {
        const double ox = o[3 * i], oy = o[3 * i + 1], oz = o[3 * i + 2];
        const double dx = d[3 * i], dy = d[3 * i + 1], dz = d[3 * i + 2];
        double best = INFINITY;
        if (dz < 0) {
            double t = -oz / dz;
            if (t > 0) best = t;
        }
        const double idx = 1.0 / dx, idy = 1.0 / dy, idz = 1.0 / dz;
        for (size_t b = 0; b < nb; ++b) {
            const double *B = boxes + 7 * b;
            if (B[3] < ox - max_range || B[0] > ox + max_range) continue;
            double t1 = (B[0] - ox) * idx, t2 = (B[3] - ox) * idx;
            double tmin = fmin(t1, t2), tmax = fmax(t1, t2);
            t1 = (B[1] - oy) * idy; t2 = (B[4] - oy) * idy;
            tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
            t1 = (B[2] - oz) * idz; t2 = (B[5] - oz) * idz;
            tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
            if (!(tmax >= tmin) || tmin <= 0 || tmin >= best) continue;
            if (B[6] > 0) {
                double u = u01(splitmix64(seed ^ splitmix64((uint64_t) i * 1315423911ull + b)));
                double t = tmin - B[6] * log(u);
                if (t > tmax || t >= best) continue;
                best = t;
            } else
                best = tmin;
        }
        for (size_t c = 0; c < nc; ++c) {
            const double *Cy = cyl + 4 * c;
            if (fabs(Cy[0] - ox) > max_range) continue;
            double px = ox - Cy[0], py = oy - Cy[1];
            double a = dx * dx + dy * dy;
            if (a < 1e-12) continue;
            double bq = 2 * (px * dx + py * dy), cq = px * px + py * py - Cy[2] * Cy[2];
            double disc = bq * bq - 4 * a * cq;
            if (disc <= 0) continue;
            double t = (-bq - sqrt(disc)) / (2 * a);
            if (t <= 0 || t >= best) continue;
            double z = oz + t * dz;
            if (z < 0 || z > Cy[3]) continue;
            best = t;
        }
        out[i] = best;
    }
}
