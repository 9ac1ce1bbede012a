/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the height-scan raycast the reference delegates to
 * third-party code that is ABSENT from /root/reference:
 *
 *   ORBIT  omni.isaac.orbit.utils.warp.ops.raycast_mesh / kernels.raycast_mesh_kernel
 *          (github.com/NVIDIA-Omniverse/orbit, unpinned, ~Feb-Mar 2024)
 *   warp   wp.mesh_query_ray (native/mesh.h, bvh.h, intersect.h; the build bundled with
 *          Isaac Sim 2023.1.x, docker/.env:4)
 *
 * Reference call sites that anchor the semantics: rover_env_cfg.py:78-86 (sensor wiring),
 * observations.py:42-45 (consumer).  Published algorithm restated:
 * closest hit along the ray with 0 <= t < max_t over a triangle mesh, found by a bounding-volume
 * hierarchy walked with a per-ray stack (AABB slab test) and the watertight ray/triangle test of
 * Woop, Benthin & Wald, "Watertight Ray/Triangle Intersection", JCGT 2(1), 2013 (double-sided,
 * double-precision fallback when an edge function is exactly zero).  A hit writes start + t*dir;
 * a miss leaves +inf.
 *
 * PARITY UNPINNED: the reference ships no test or golden vector for the raycast;
 * this file is pinned instead by (i) a brute-force all-triangles mode that must agree with the BVH
 * mode bit for bit (tests/test_host_cpu.py::test_oracle_bvh_equals_brute_force), and (ii) analytic surfaces derived
 * independently in float64 -- heightfield interpolation, tilted planes, misses --
 * (tests/test_host_cpu.py::test_oracle_raycast_against_independent_float64_geometry).
 *
 * Build: oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).  Only tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the resulting library.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float lo[3], hi[3];
    int left;   /* internal: index of left child (right = left + 1); leaf: first primitive */
    int count;  /* 0 for internal nodes, else number of primitives */
} Node;

typedef struct {
    int nv, nf;
    float *v;      /* nv x 3 */
    int *f;        /* nf x 3 */
    int *prim;     /* permutation of faces referenced by leaves */
    Node *nodes;
    int n_nodes;
} Mesh;

#define LEAF_SIZE 4

static void face_bounds(const Mesh *m, int fi, float lo[3], float hi[3]) {
    const int *t = m->f + 3 * fi;
    for (int k = 0; k < 3; ++k) {
        float a = m->v[3 * t[0] + k], b = m->v[3 * t[1] + k], c = m->v[3 * t[2] + k];
        lo[k] = fminf(a, fminf(b, c));
        hi[k] = fmaxf(a, fmaxf(b, c));
    }
}

typedef struct { const float *cent; int axis; } SortCtx;
static SortCtx g_ctx; /* build is single-threaded */
static int cmp_centroid(const void *pa, const void *pb) {
    float a = g_ctx.cent[3 * (*(const int *)pa) + g_ctx.axis];
    float b = g_ctx.cent[3 * (*(const int *)pb) + g_ctx.axis];
    return (a > b) - (a < b);
}

/* median split on the widest centroid axis */
static void build_node(Mesh *m, const float *cent, int node, int first, int count) {
    Node *nd = &m->nodes[node];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = first; i < first + count; ++i) {
        float l[3], h[3];
        face_bounds(m, m->prim[i], l, h);
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], l[k]);
            hi[k] = fmaxf(hi[k], h[k]);
            float c = cent[3 * m->prim[i] + k];
            clo[k] = fminf(clo[k], c);
            chi[k] = fmaxf(chi[k], c);
        }
    }
    memcpy(nd->lo, lo, sizeof lo);
    memcpy(nd->hi, hi, sizeof hi);
    if (count <= LEAF_SIZE) {
        nd->left = first;
        nd->count = count;
        return;
    }
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    g_ctx.cent = cent;
    g_ctx.axis = axis;
    qsort(m->prim + first, (size_t)count, sizeof(int), cmp_centroid);
    int half = count / 2;
    int l = m->n_nodes;
    m->n_nodes += 2;
    nd = &m->nodes[node];
    nd->left = l;
    nd->count = 0;
    build_node(m, cent, l, first, half);
    build_node(m, cent, l + 1, first + half, count - half);
}

Mesh *rc_mesh_create(const float *verts, int nv, const int *faces, int nf) {
    Mesh *m = (Mesh *)calloc(1, sizeof(Mesh));
    m->nv = nv;
    m->nf = nf;
    m->v = (float *)malloc(sizeof(float) * 3 * (size_t)nv);
    m->f = (int *)malloc(sizeof(int) * 3 * (size_t)nf);
    memcpy(m->v, verts, sizeof(float) * 3 * (size_t)nv);
    memcpy(m->f, faces, sizeof(int) * 3 * (size_t)nf);
    m->prim = (int *)malloc(sizeof(int) * (size_t)(nf > 0 ? nf : 1));
    m->nodes = (Node *)malloc(sizeof(Node) * (size_t)(2 * (nf > 0 ? nf : 1)));
    float *cent = (float *)malloc(sizeof(float) * 3 * (size_t)(nf > 0 ? nf : 1));
    for (int i = 0; i < nf; ++i) {
        m->prim[i] = i;
        float l[3], h[3];
        face_bounds(m, i, l, h);
        for (int k = 0; k < 3; ++k) cent[3 * i + k] = 0.5f * (l[k] + h[k]);
    }
    m->n_nodes = 1;
    if (nf > 0) build_node(m, cent, 0, 0, nf);
    else { m->nodes[0].left = 0; m->nodes[0].count = 0; for (int k = 0; k < 3; ++k) { m->nodes[0].lo[k] = FLT_MAX; m->nodes[0].hi[k] = -FLT_MAX; } }
    free(cent);
    return m;
}

void rc_mesh_destroy(Mesh *m) {
    if (!m) return;
    free(m->v); free(m->f); free(m->prim); free(m->nodes); free(m);
}

static inline int max_dim(const float d[3]) {
    float x = fabsf(d[0]), y = fabsf(d[1]), z = fabsf(d[2]);
    return x > y ? (x > z ? 0 : 2) : (y > z ? 1 : 2);
}

/* Woop-Benthin-Wald watertight test; returns 1 and *t_out on a hit with t >= 0 (sign-corrected). */
static int ray_tri_watertight(const float org[3], const float dir[3], const float *a, const float *b, const float *c,
                              float *t_out) {
    int kz = max_dim(dir);
    int kx = (kz + 1) % 3;
    int ky = (kx + 1) % 3;
    if (dir[kz] < 0.0f) { int tmp = kx; kx = ky; ky = tmp; }
    const float Sx = dir[kx] / dir[kz];
    const float Sy = dir[ky] / dir[kz];
    const float Sz = 1.0f / dir[kz];
    const float A[3] = {a[0] - org[0], a[1] - org[1], a[2] - org[2]};
    const float B[3] = {b[0] - org[0], b[1] - org[1], b[2] - org[2]};
    const float C[3] = {c[0] - org[0], c[1] - org[1], c[2] - org[2]};
    const float Ax = A[kx] - Sx * A[kz], Ay = A[ky] - Sy * A[kz];
    const float Bx = B[kx] - Sx * B[kz], By = B[ky] - Sy * B[kz];
    const float Cx = C[kx] - Sx * C[kz], Cy = C[ky] - Sy * C[kz];
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
    const float det = U + V + W;
    if (det == 0.0f) return 0;
    const float Az = Sz * A[kz], Bz = Sz * B[kz], Cz = Sz * C[kz];
    const float T = U * Az + V * Bz + W * Cz;
    /* t = T / det must be >= 0: T and det of the same sign (double-sided) */
    if ((det < 0.0f && T > 0.0f) || (det > 0.0f && T < 0.0f)) return 0;
    const float rcp = 1.0f / det;
    *t_out = T * rcp;
    return 1;
}

static inline int ray_aabb(const float org[3], const float dir[3], const float rcp[3], const float lo[3],
                           const float hi[3], float tmax) {
    float t0 = 0.0f, t1 = tmax;
    for (int k = 0; k < 3; ++k) {
        if (dir[k] == 0.0f) {
            /* axis-parallel ray: inside the (closed) slab or not; avoids the 0 * inf = NaN of the slab form */
            if (org[k] < lo[k] || org[k] > hi[k]) return 0;
            continue;
        }
        float a = (lo[k] - org[k]) * rcp[k];
        float b = (hi[k] - org[k]) * rcp[k];
        float n = fminf(a, b), f = fmaxf(a, b);
        if (n > t0) t0 = n;
        if (f < t1) t1 = f;
    }
    /* widen by a few ulp so that rays grazing a box face are never culled */
    return t0 <= t1 * 1.00001f + 1e-6f;
}

static void cast_one(const Mesh *m, const float org[3], const float dir[3], float max_t, int brute, float *t_hit,
                     int *f_hit) {
    float best = max_t;
    int best_f = -1;
    if (brute) {
        for (int i = 0; i < m->nf; ++i) {
            const int *t = m->f + 3 * i;
            float tt;
            if (ray_tri_watertight(org, dir, m->v + 3 * t[0], m->v + 3 * t[1], m->v + 3 * t[2], &tt) && tt < best &&
                tt >= 0.0f) {
                best = tt;
                best_f = i;
            }
        }
    } else if (m->nf > 0) {
        float rcp[3] = {1.0f / dir[0], 1.0f / dir[1], 1.0f / dir[2]};
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const Node *nd = &m->nodes[stack[--sp]];
            if (!ray_aabb(org, dir, rcp, nd->lo, nd->hi, best)) continue;
            if (nd->count > 0) {
                for (int i = nd->left; i < nd->left + nd->count; ++i) {
                    int fi = m->prim[i];
                    const int *t = m->f + 3 * fi;
                    float tt;
                    /* ties on t keep the lowest face index, like the in-order brute-force scan */
                    if (ray_tri_watertight(org, dir, m->v + 3 * t[0], m->v + 3 * t[1], m->v + 3 * t[2], &tt) &&
                        tt >= 0.0f && (tt < best || (tt == best && best_f >= 0 && fi < best_f))) {
                        best = tt;
                        best_f = fi;
                    }
                }
            } else if (sp + 2 <= 128) {
                stack[sp++] = nd->left;
                stack[sp++] = nd->left + 1;
            }
        }
    }
    *t_hit = best;
    *f_hit = best_f;
}

/* starts, dirs: n x 3;  hits: n x 3 (written: start + t*dir, or +inf on a miss);  t_out / face_out optional */
void rc_raycast(const Mesh *m, const float *starts, const float *dirs, int64_t n, float max_t, int brute, float *hits,
                float *t_out, int *face_out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        const float *o = starts + 3 * i, *d = dirs + 3 * i;
        float t;
        int f;
        cast_one(m, o, d, max_t, brute, &t, &f);
        if (f >= 0) {
            hits[3 * i + 0] = o[0] + t * d[0];
            hits[3 * i + 1] = o[1] + t * d[1];
            hits[3 * i + 2] = o[2] + t * d[2];
        } else {
            hits[3 * i + 0] = hits[3 * i + 1] = hits[3 * i + 2] = INFINITY;
            t = INFINITY;
        }
        if (t_out) t_out[i] = t;
        if (face_out) face_out[i] = f;
    }
}

int rc_num_nodes(const Mesh *m) { return m->n_nodes; }
