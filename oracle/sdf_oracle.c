/* ORACLE (test infrastructure, not product code) — CPU restatement of the voxel-SDF kernel.
 *
 * PARITY UNPINNED.  The reference computes its interpenetration field in the un-vendored,
 * un-pinned CUDA extension `sdf` (github.com/penincillin/SDF_ihmr, installed from git HEAD,
 * /root/reference/docs/install.md:37; call site /root/reference/src/models/loss_utils.py:38,
 * 181-182).  Its source is not available offline, so this file restates the published
 * algorithm of that package's `sdf_cuda_kernel.cu` lineage (JiangWenPL/multiperson/sdf)
 * under assumptions A1, A3, A4 of SURVEY.md §8(c): for every voxel centre of a G^3 grid over
 * [-1,1]^3, loop over ALL faces, take the minimum point-triangle distance and the parity of
 * +x ray crossings; phi = distance if inside (odd parity) else 0.  One voxel per loop
 * iteration, brute force, exactly like the one-thread-per-voxel reference kernel.
 *
 * Arithmetic contract shared with the CUDA path (so inside/outside decisions agree bit for
 * bit on identical inputs): no FMA contraction (compile with -ffp-contract=off), the edge
 * functions are evaluated on the (lower vertex id, higher vertex id) ordering of each edge so
 * two faces sharing an edge see the same value, ties go to the face that walks the edge in
 * ascending id order.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define SQRT(x) ((REAL)sqrt((double)(x)))

static inline int edge_side(const REAL* P, int i0, int i1, REAL qy, REAL qz, REAL* w) {
    int fwd = i0 < i1;
    const REAL* lo = P + 3 * (fwd ? i0 : i1);
    const REAL* hi = P + 3 * (fwd ? i1 : i0);
    REAL e = (hi[1] - lo[1]) * (qz - lo[2]) - (hi[2] - lo[2]) * (qy - lo[1]);
    *w = fwd ? e : -e;
    return fwd ? (e >= 0) : (e < 0);
}

/* 1 if the +x ray from q crosses face (ia, ib, ic) strictly beyond q.x */
static inline int ray_cross(const REAL* P, int ia, int ib, int ic, const REAL* q) {
    REAL wa, wb, wc;
    int p0 = edge_side(P, ia, ib, q[1], q[2], &wc);
    int p1 = edge_side(P, ib, ic, q[1], q[2], &wa);
    int p2 = edge_side(P, ic, ia, q[1], q[2], &wb);
    if (!((p0 && p1 && p2) || (!p0 && !p1 && !p2))) return 0;
    REAL sum = (wa + wb) + wc;
    if (sum == 0) return 0;
    REAL x = ((wa * P[3 * ia] + wb * P[3 * ib]) + wc * P[3 * ic]) / sum;
    return x > q[0];
}

static inline REAL dot3(const REAL* a, const REAL* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* squared distance from p to triangle (a, b, c): closest-point regions (vertex / edge / face) */
static inline REAL pt_tri_dist2(const REAL* p, const REAL* a, const REAL* b, const REAL* c) {
    REAL ab[3], ac[3], ap[3], bp[3], cp[3], cl[3], d[3];
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    REAL d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) return dot3(ap, ap);
    for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
    REAL d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) return dot3(bp, bp);
    for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
    REAL d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) return dot3(cp, cp);
    REAL vc = d1 * d4 - d3 * d2;
    REAL vb = d5 * d2 - d1 * d6;
    REAL va = d3 * d6 - d5 * d4;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        REAL t = d1 / (d1 - d3);
        for (int k = 0; k < 3; ++k) cl[k] = a[k] + t * ab[k];
    } else if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        REAL t = d2 / (d2 - d6);
        for (int k = 0; k < 3; ++k) cl[k] = a[k] + t * ac[k];
    } else if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        REAL t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; ++k) cl[k] = b[k] + t * (c[k] - b[k]);
    } else {
        REAL den = (va + vb) + vc;
        if (den == 0) {  /* degenerate face: nearest corner */
            REAL m = dot3(ap, ap), m2 = dot3(bp, bp), m3 = dot3(cp, cp);
            m = m2 < m ? m2 : m;
            return m3 < m ? m3 : m;
        }
        REAL v = vb / den, w = vc / den;
        for (int k = 0; k < 3; ++k) cl[k] = (a[k] + v * ab[k]) + w * ac[k];
    }
    for (int k = 0; k < 3; ++k) d[k] = p[k] - cl[k];
    return dot3(d, d);
}

/* phi (n_mesh, G, G, G) laid out [z][y][x] (A3); verts (n_mesh, nv, 3) already normalised
 * into [-1,1]^3; faces (nf, 3) int32 shared by all meshes. */
void CAT(sdf_grid, SUFFIX)(const REAL* verts, const int32_t* faces, int n_mesh, int nv, int nf,
                           int G, REAL* phi) {
    long total = (long)n_mesh * G * G * G;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < total; ++i) {
        int m = (int)(i / ((long)G * G * G));
        int pn = (int)(i % ((long)G * G * G));
        int zi = pn / (G * G), yi = (pn / G) % G, xi = pn % G;
        REAL q[3];
        q[0] = ((REAL)2 * xi + 1 - G) / G;
        q[1] = ((REAL)2 * yi + 1 - G) / G;
        q[2] = ((REAL)2 * zi + 1 - G) / G;
        const REAL* P = verts + (long)m * nv * 3;
        REAL best = (REAL)1e30;
        int crossings = 0;
        for (int f = 0; f < nf; ++f) {
            int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
            crossings += ray_cross(P, ia, ib, ic, q);
            REAL d2 = pt_tri_dist2(q, P + 3 * ia, P + 3 * ib, P + 3 * ic);
            best = d2 < best ? d2 : best;
        }
        phi[i] = (crossings & 1) ? SQRT(best) : (REAL)0;
    }
}

/* single-voxel probes used by the unit tests */
int CAT(sdf_ray_cross, SUFFIX)(const REAL* verts, const int32_t* face, const REAL* q) {
    return ray_cross(verts, face[0], face[1], face[2], q);
}
REAL CAT(sdf_pt_tri_dist2, SUFFIX)(const REAL* p, const REAL* a, const REAL* b, const REAL* c) {
    return pt_tri_dist2(p, a, b, c);
}
