/*
 * pfo_kd.cpp -- CPU ORACLE, kd-tree point-cloud path (test infrastructure only; see pfo.h).
 *
 * Restates the kd variants of the reference's step (src/kernel.cu:816-1540, :1702-1761) and its host
 * kd-tree code (src/kdtree.cpp:25-123).  C++ rather than C for one reason: KDTree::Create/Balance
 * sort with std::sort, whose order among equal keys (every third level sorts on z == 0, and x/y are
 * multiples of the cell size) decides the tree shape; calling the same libstdc++ std::sort on the same
 * sequence reproduces the reference build exactly (pinned against oracle/_ref in
 * tests/test_oracle_kd.py).
 *
 * Definitions where the reference reads uninitialised or out-of-bounds memory:
 *   Q9  the NN walk stops when the best node is the root (parent == -1) instead of reading tree[-1];
 *   Q10 only the first |wallPC| free points are valid on the device (kernel.cu:1475 copies the wrong
 *       count): exactly those are applied, the uninitialised rest is skipped;
 *   Q11 ICP targets of beams that fail the +-20 m filter are (0,0,0) (never written by kernGetWallsKD).
 * Sums whose association the reference leaves to thrust::reduce (ICP means and W) use the fixed
 * "ICP order": 32 strided sequential partial sums, then an xor-butterfly (16,8,4,2,1).
 * The 3x3 SVD (svd3.h) is replaced by the closed form of the planar case it is used for (z == 0):
 * R = U V^T is the polar rotation of the 2x2 block, angle atan2(H01 - H10, H00 + H11).
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pfo.h"

extern "C" {

typedef struct { int axis, left, right, parent; float x, y, z, w; } pfo_kdnode;   /* == KDTree::Node, 32 B */

struct V4 { float x, y, z, w; };
static bool lessX(const V4 &a, const V4 &b) { return a.x < b.x; }   /* kdtree.cpp:12-23 */
static bool lessY(const V4 &a, const V4 &b) { return a.y < b.y; }
static bool lessZ(const V4 &a, const V4 &b) { return a.z < b.z; }

static void set_node(pfo_kdnode *n, const V4 &p, int axis, int parent)     /* kdtree.cpp:116-122 */
{
    n->left = -1; n->right = -1; n->parent = parent; n->axis = axis;
    n->x = p.x; n->y = p.y; n->z = p.z; n->w = p.w;
}

/* kdtree.cpp:42-67 InsertList; sorting the sub-range in place == sorting the reference's copy */
static void insert_list(V4 *first, V4 *last, pfo_kdnode *list, int idx, int parent)
{
    int axis = parent == -1 ? 0 : (list[parent].axis + 1) % 3;
    if (axis == 0) std::sort(first, last, lessX);
    if (axis == 1) std::sort(first, last, lessY);
    if (axis == 2) std::sort(first, last, lessZ);
    const int size = (int)(last - first);
    const int mid = size / 2;
    set_node(&list[idx], first[mid], axis, parent);
    if (mid > 0) { list[idx].left = idx + 1; insert_list(first, first + mid, list, idx + 1, idx); }
    if (mid < size - 1) { list[idx].right = idx + mid + 1; insert_list(first + mid + 1, last, list, idx + mid + 1, idx); }
}

/* kdtree.cpp:25-29 Create */
void pfo_kd_create(const float *pts4, int n, pfo_kdnode *list)
{
    if (n <= 0) return;
    std::vector<V4> v(n);
    memcpy(v.data(), pts4, (size_t)n * 16);
    std::sort(v.begin(), v.end(), lessX);
    insert_list(v.data(), v.data() + n, list, 0, -1);
}

/* kdtree.cpp:31-40 Balance */
void pfo_kd_balance(pfo_kdnode *list, int size)
{
    std::vector<float> pts((size_t)size * 4);
    for (int i = 0; i < size; i++) { pts[4 * i] = list[i].x; pts[4 * i + 1] = list[i].y; pts[4 * i + 2] = list[i].z; pts[4 * i + 3] = list[i].w; }
    pfo_kd_create(pts.data(), size, list);
}

/* kdtree.cpp:69-105 InsertNode */
void pfo_kd_insert(const float *pt4, pfo_kdnode *list, int size)
{
    V4 p = {pt4[0], pt4[1], pt4[2], pt4[3]};
    int next = 0, parent = 0, axis = 0;
    bool less = false;
    do {
        parent = next;
        axis = list[next].parent == -1 ? 0 : (list[list[next].parent].axis + 1) % 3;
        const float a = axis == 0 ? p.x : axis == 1 ? p.y : p.z;
        const float b = axis == 0 ? list[next].x : axis == 1 ? list[next].y : list[next].z;
        less = a < b;
        next = less ? list[next].left : list[next].right;
    } while (next != -1);
    if (less) list[parent].left = size; else list[parent].right = size;
    set_node(&list[size], p, (axis + 1) % 3, parent);
}

/* glm::distance as the device code evaluates it (SASS): sqrt(fma(dz,dz, fma(dx,dx, dy*dy))) */
static inline float dist3(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = bx - ax, dy = by - ay, dz = bz - az;
    return sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
}

/* kernel.cu:924-972 findCorrespondenceIndexKD (== :874-922, :1147-1184, :1239-1276) */
int pfo_kd_nn(const pfo_kdnode *tree, float qx, float qy, float qz)
{
    float bestDist = dist3(qx, qy, qz, tree[0].x, tree[0].y, tree[0].z);
    int bestIdx = 0, head = 0;
    bool done = false, explored = false;
    while (!done) {
        while (head >= 0) {
            const pfo_kdnode &t = tree[head];
            const float d = dist3(qx, qy, qz, t.x, t.y, t.z);
            if (d < bestDist) { bestDist = d; bestIdx = head; explored = false; }
            const bool branch = t.axis == 0 ? qx < t.x : t.axis == 1 ? qy < t.y : t.axis == 2 ? qz < t.z : false;
            head = branch ? t.left : t.right;
        }
        if (explored) done = true;
        else {
            const int pi = tree[bestIdx].parent;
            if (pi < 0) { done = true; break; }                               /* Q9 */
            const pfo_kdnode &p = tree[pi];
            bool branch = false; float hd = 0.0f;
            if (p.axis == 0) { branch = qx < p.x; hd = fabsf(qx - p.x); }
            if (p.axis == 1) { branch = qy < p.y; hd = fabsf(qy - p.y); }
            if (p.axis == 2) { branch = qz < p.z; hd = fabsf(qz - p.z); }
            if (hd < bestDist) { head = !branch ? p.left : p.right; explored = true; }
            else done = true;
        }
    }
    return bestIdx;
}

/* kernel.cu:1198-1298 EvaluateParticleKD (DYNAMIC_KERN 0); node weights are integers, so is the sum */
int pfo_kd_score(const pfo_config *c, const pfo_kdnode *tree, float px, float py, float pth, const float *scan)
{
    int sum = 0;
    for (int j = 0; j < c->n_beams; j++) {
        const float rot = pfo_lidar_angle(j) + pth;
        const float wx = scan[j] * (c->trig == PFO_TRIG_CUDA ? pfo_cosf_cuda(rot) : cosf(rot));
        const float wy = scan[j] * (c->trig == PFO_TRIG_CUDA ? pfo_sinf_cuda(rot) : sinf(rot));
        if (fabsf(wx) < PFO_LIDAR_RANGE && fabsf(wy) < PFO_LIDAR_RANGE)
            sum += (int)tree[pfo_kd_nn(tree, wx + px, wy + py, 0.0f)].w;
    }
    return sum;
}

/* IEEE-only asin (Cephes asinf scheme) for theta = asin(R[0][1]), kernel.cu:1079 */
float pfo_asinf(float x)
{
    float a = fabsf(x);
    if (a > 1.0f) a = 1.0f;
    int flag = 0; float z, xx;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); xx = sqrtf(z); flag = 1; }
    else { xx = a; z = xx * xx; }
    float p = 4.2163199048E-2f;
    p = fmaf(p, z, 2.4181311049E-2f);
    p = fmaf(p, z, 4.5470025998E-2f);
    p = fmaf(p, z, 7.4953002686E-2f);
    p = fmaf(p, z, 1.6666752422E-1f);
    float r = fmaf(p * z, xx, xx);
    if (flag) { r = r + r; r = 1.570796326794896619f - r; }
    return x < 0.0f ? -r : r;
}

/* fixed "ICP order" sum of n floats with stride */
static float icp_sum(const float *v, int n, int stride)
{
    float part[32];
    for (int l = 0; l < 32; l++) { float s = 0.0f; for (int i = l; i < n; i += 32) s = s + v[(size_t)i * stride]; part[l] = s; }
    for (int o = 16; o > 0; o >>= 1) { float nw[32]; for (int l = 0; l < 32; l++) nw[l] = part[l] + part[l ^ o]; memcpy(part, nw, sizeof nw); }
    return part[0];
}

/* kernel.cu:993-1093 transformPointICP: one point-to-point step; targets from the PREVIOUS robotPos */
void pfo_kd_icp(const pfo_config *c, const pfo_kdnode *tree, const float robot_prev[3], const float start[3],
                const float *scan, float out[3])
{
    const int B = c->n_beams;
    std::vector<float> tar((size_t)B * 3, 0.0f), cor((size_t)B * 3, 0.0f);
    for (int i = 0; i < B; i++) {
        const float rot = pfo_lidar_angle(i) + robot_prev[2];
        const float wx = scan[i] * (c->trig == PFO_TRIG_CUDA ? pfo_cosf_cuda(rot) : cosf(rot));
        const float wy = scan[i] * (c->trig == PFO_TRIG_CUDA ? pfo_sinf_cuda(rot) : sinf(rot));
        if (fabsf(wx) < PFO_LIDAR_RANGE && fabsf(wy) < PFO_LIDAR_RANGE) { tar[3 * i] = robot_prev[0] + wx; tar[3 * i + 1] = robot_prev[1] + wy; }
        const pfo_kdnode &nn = tree[pfo_kd_nn(tree, tar[3 * i], tar[3 * i + 1], tar[3 * i + 2])];
        cor[3 * i] = nn.x; cor[3 * i + 1] = nn.y; cor[3 * i + 2] = nn.z;
    }
    float mt[2], mc[2];
    for (int k = 0; k < 2; k++) { mt[k] = icp_sum(&tar[k], B, 3) / (float)B; mc[k] = icp_sum(&cor[k], B, 3) / (float)B; }
    std::vector<float> prod((size_t)B);
    float H[2][2];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            for (int q = 0; q < B; q++) prod[q] = (tar[3 * q + i] - mt[i]) * (cor[3 * q + j] - mc[j]);
            H[i][j] = icp_sum(prod.data(), B, 1);
        }
    const float s = H[0][1] - H[1][0], k = H[0][0] + H[1][1];
    const float nrm = sqrtf(fmaf(s, s, k * k));
    float sn = 0.0f, cs = 1.0f;
    if (nrm > 0.0f) { sn = s / nrm; cs = k / nrm; }
    const float tx = mc[0] - fmaf(cs, mt[0], -(sn * mt[1]));
    const float ty = mc[1] - fmaf(sn, mt[0], cs * mt[1]);
    out[0] = start[0] + tx; out[1] = start[1] + ty; out[2] = start[2] + pfo_asinf(sn);
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    pfo_config cfg;
    int n;
    float *x, *y, *th, *w, *weff;
    int32_t *fit;
    float *cdf;
    pfo_kdnode *tree; int kd_size, kd_cap;
    uint8_t *free_mask, *wall_mask;
    float robot[3];
    int32_t fit_min, fit_max; int best;
    float sum_w, sum_w2, neff; int resampled;
    int n_free_pts, n_wall_pts, n_inserted;
} pfo_kd_state;

pfo_kd_state *pfo_kd_create_state(const pfo_config *c, int n, int kd_cap)
{
    pfo_kd_state *s = (pfo_kd_state *)calloc(1, sizeof *s);
    s->cfg = *c; s->n = n; s->kd_cap = kd_cap;
    s->x = (float *)calloc(n, 4); s->y = (float *)calloc(n, 4); s->th = (float *)calloc(n, 4);
    s->w = (float *)malloc((size_t)n * 4); s->weff = (float *)malloc((size_t)n * 4);
    for (int i = 0; i < n; i++) { s->w[i] = 1.0f; s->weff[i] = 1.0f; }
    s->fit = (int32_t *)calloc(n, 4); s->cdf = (float *)calloc(n, 4);
    s->tree = (pfo_kdnode *)calloc((size_t)kd_cap, sizeof(pfo_kdnode));
    size_t nc = (size_t)c->map_w * c->map_h;
    s->free_mask = (uint8_t *)calloc(nc, 1); s->wall_mask = (uint8_t *)calloc(nc, 1);
    return s;
}

void pfo_kd_destroy_state(pfo_kd_state *s)
{
    if (!s) return;
    free(s->x); free(s->y); free(s->th); free(s->w); free(s->weff); free(s->fit); free(s->cdf);
    free(s->tree); free(s->free_mask); free(s->wall_mask); free(s);
}

static inline float round_frac(float a, float frac) { return roundf(a / frac) * frac; }   /* kernel.cu:52 */

/* kernel.cu:1406-1540 PFUpdateMapKD; hits_free / hits_wall (optional, kd_size ints each, zeroed by the caller):
 * points within range per node in the free / wall weight pass */
void pfo_kd_update_map_hits(pfo_kd_state *s, const float *scan, int *hits_free, int *hits_wall);
void pfo_kd_update_map(pfo_kd_state *s, const float *scan) { pfo_kd_update_map_hits(s, scan, NULL, NULL); }

void pfo_kd_update_map_hits(pfo_kd_state *s, const float *scan, int *hits_free, int *hits_wall)
{
    const pfo_config *c = &s->cfg;
    const size_t nc = (size_t)c->map_w * c->map_h;
    const int cx = (int)roundf(0.5f * (float)c->map_w + c->res_x / 2.0f);               /* kernel.cu:1408-1411 */
    const int cy = (int)roundf(0.5f * (float)c->map_h + c->res_y / 2.0f);
    memset(s->free_mask, 0, nc); memset(s->wall_mask, 0, nc);
    pfo_get_walls(c, scan, cx, cy, s->robot[2], s->free_mask, s->wall_mask);
    std::vector<V4> wallPC, freePC;
    for (int x = 0; x < c->map_w; x++)
        for (int y = 0; y < c->map_h; y++) {
            const size_t idx = (size_t)x * c->map_w + y;
            if (!s->wall_mask[idx] && !s->free_mask[idx]) continue;
            V4 p;
            float px = (float)x * c->res_x; px = px - c->scale_x / 2.0f; px = px + s->robot[0];
            float py = (float)y * c->res_y; py = py - c->scale_y / 2.0f; py = py + s->robot[1];
            p.x = round_frac(px, c->res_x); p.y = round_frac(py, c->res_y); p.z = 0.0f; p.w = 0.0f;
            if (s->wall_mask[idx]) wallPC.push_back(p);
            if (s->free_mask[idx]) freePC.push_back(p);
        }
    s->n_wall_pts = (int)wallPC.size(); s->n_free_pts = (int)freePC.size(); s->n_inserted = 0;
    if (s->kd_size > 0) {
        const int nW = (int)wallPC.size();
        const int nF = std::min((int)freePC.size(), nW);                                 /* Q10 */
        std::vector<int> iF(nF), iW(nW);
        for (int i = 0; i < nF; i++) iF[i] = pfo_kd_nn(s->tree, freePC[i].x, freePC[i].y, freePC[i].z);
        for (int i = 0; i < nW; i++) iW[i] = pfo_kd_nn(s->tree, wallPC[i].x, wallPC[i].y, wallPC[i].z);
        const float minDist = sqrtf(fmaf(c->res_y, c->res_y, c->res_x * c->res_x));      /* kernel.cu:1358 */
        for (int pass = 0; pass < 2; pass++) {                                            /* kernel.cu:1492-1495 */
            const std::vector<V4> &pc = pass == 0 ? freePC : wallPC;
            const std::vector<int> &ix = pass == 0 ? iF : iW;
            const float val = pass == 0 ? (float)PFO_FREE_WEIGHT : (float)PFO_OCCUPIED_WEIGHT;
            /* kernUpdateMapKD (kernel.cu:1350-1364) updates tree[idx].value.w with a plain load and store from
             * one thread per point.  Points that share a nearest node therefore RACE in the reference: of k
             * colliding threads, any number between 1 and k may take effect, depending on which of them read
             * the weight before the others stored it (lanes of one warp always collapse to one; warps that
             * overlap in time collapse, warps that do not accumulate).  On the B200 the reference's kernel
             * shows both outcomes on the same frame (T3, test_kd_map_update_equals_reference).  The engine and
             * this oracle define the deterministic outcome "ONCE per node per launch" -- always one of the
             * race's legal results, and the kd counterpart of the grid path's once-per-cell bool masks
             * (kernel.cu:513-522).  `hits` (optional) receives the number of colliding points per node so
             * that a test can check the reference's value against the legal range. */
            std::vector<char> touched((size_t)s->kd_size, 0);
            int *hits = pass == 0 ? hits_free : hits_wall;
            for (size_t i = 0; i < ix.size(); i++) {
                pfo_kdnode &t = s->tree[ix[i]];
                if (dist3(pc[i].x, pc[i].y, pc[i].z, t.x, t.y, t.z) < minDist) {
                    if (hits) hits[ix[i]]++;
                    if (touched[ix[i]]) continue;
                    touched[ix[i]] = 1;
                    float v = t.w + val;
                    t.w = v < -(float)PFO_CLAMP_VAL ? -(float)PFO_CLAMP_VAL : v > (float)PFO_CLAMP_VAL ? (float)PFO_CLAMP_VAL : v;
                }
            }
        }
        std::vector<char> diff(nW);
        for (int i = 0; i < nW; i++) {                                                    /* kernel.cu:1367-1379 */
            const pfo_kdnode &t = s->tree[iW[i]];
            diff[i] = dist3(wallPC[i].x, wallPC[i].y, wallPC[i].z, t.x, t.y, t.z) > minDist * 0.5f;
        }
        for (int i = 0; i < nW; i++)                                                      /* kernel.cu:1512-1517 */
            if (diff[i] && s->kd_size < s->kd_cap) {
                float p4[4] = {wallPC[i].x, wallPC[i].y, wallPC[i].z, -100.0f};
                pfo_kd_insert(p4, s->tree, s->kd_size++);
                s->n_inserted++;
            }
    } else if (!wallPC.empty()) {                                                         /* kernel.cu:1532-1536 */
        pfo_kd_create(&wallPC[0].x, (int)wallPC.size(), s->tree);
        s->kd_size = (int)wallPC.size();
        s->n_inserted = s->kd_size;
    }
}

/* kernel.cu:1311-1348 PFMeasurementUpdateKD */
void pfo_kd_measure(pfo_kd_state *s, const float *scan)
{
    for (int i = 0; i < s->n; i++) s->fit[i] = pfo_kd_score(&s->cfg, s->tree, s->x[i], s->y[i], s->th[i], scan);
    pfo_minmax(s->fit, s->n, &s->fit_min, &s->fit_max, &s->best);
    const int rng = s->fit_max - s->fit_min;
    if (rng > 0) {
        const float f = 1.0f / (float)rng, fmin = (float)s->fit_min;                      /* Q5: (int)min is exact here */
        for (int i = 0; i < s->n; i++) s->weff[i] = (s->weff[i] * ((float)s->fit[i] - fmin)) * f;
    }
    const int n_sync = s->cfg.quirk_q1 ? (s->n + 1) / 2 : s->n;                          /* kernel.cu:1341 */
    memcpy(s->w, s->weff, (size_t)n_sync * 4);
    const float start[3] = {s->x[s->best], s->y[s->best], s->th[s->best]};
    float out[3];
    pfo_kd_icp(&s->cfg, s->tree, s->robot, start, scan, out);                            /* kernel.cu:1345 */
    s->robot[0] = out[0]; s->robot[1] = out[1]; s->robot[2] = out[2];
}

/* kernel.cu:1702-1761 particleFilter (kd variant) */
void pfo_kd_step(pfo_kd_state *s, const float *scan, int frame)
{
    if (frame % 100 == 5 && s->kd_size > 0) pfo_kd_balance(s->tree, s->kd_size);          /* kernel.cu:1707-1711 */
    s->resampled = 0;
    if (s->kd_size == 0) {                                                                /* kernel.cu:1714-1717 */
        s->robot[0] = s->robot[1] = s->robot[2] = 0.0f;
        pfo_kd_update_map(s, scan);
        return;
    }
    memcpy(s->weff, s->w, (size_t)s->n * 4);                                              /* PFMotionUpdate */
    pfo_add_noise(s->x, s->y, s->th, s->n, frame, 0);
    pfo_kd_measure(s, scan);
    pfo_kd_update_map(s, scan);
    /* PFResample (kernel.cu:447-511), same as the grid path */
    const int n = s->n;
    std::vector<float> sq(n);
    for (int i = 0; i < n; i++) sq[i] = s->weff[i] * s->weff[i];
    s->sum_w2 = pfo_scan(sq.data(), n, s->cdf);
    s->sum_w = pfo_scan(s->weff, n, s->cdf);
    s->neff = (s->sum_w * s->sum_w) / s->sum_w2;
    if ((double)s->neff < PFO_EFFECTIVE * (double)n) {
        std::vector<float> ox(s->x, s->x + n), oy(s->y, s->y + n), ot(s->th, s->th + n);
        for (int i = 0; i < n; i++) {
            const int src = pfo_resample_src(s->cdf, n, s->sum_w, s->neff, frame, i);
            s->x[i] = ox[src]; s->y[i] = oy[src]; s->th[i] = ot[src];
            s->w[i] = 1.0f; s->weff[i] = 1.0f;
        }
        s->resampled = 1;
    }
}

}  /* extern "C" */
