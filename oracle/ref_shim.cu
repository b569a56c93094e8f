/*
 * ref_shim.cu -- extern "C" access to the REFERENCE's own host functions (test infrastructure).
 *
 * Linked with the reference's unmodified translation units (compiled where they lie under
 * /root/reference/src by oracle/Makefile) into oracle/_ref/libref.so.  Nothing here restates the
 * algorithm: every function forwards to the reference symbol of the same name.  lidar.cpp cannot be
 * built (MathWorks libmat absent), so the two Lidar members it would define are stubbed here.
 */
#include <vector>
#include <string>
#include <cstring>
#include "sceneStructs.h"
#include "scene.h"
#include "lidar.h"
#include "kdtree.hpp"

/* host instantiations of the reference's __host__ __device__ functions (src/kernel.cu) */
unsigned int utilhash(unsigned int a);                                           /* kernel.cu:89  */
void CleanLidarScan(int n, const float scan, const float theta, glm::vec2 &out); /* kernel.cu:182 */
void traceRay(glm::ivec2 start, glm::ivec2 end, glm::ivec2 map_dim, bool *out);   /* kernel.cu:190 */
int mapCorrelation(int N, const MAP_TYPE *map, glm::ivec2 dim, const glm::vec2 *points); /* :243 */
int EvaluateParticle(MAP_TYPE *map, glm::ivec2 map_dim, Patch map_params, Particle &particle,
                     glm::vec3 pos, float *lidar);                               /* kernel.cu:257 */
void ParticleAddNoise(Particle &particle, int frame, int idx);                   /* kernel.cu:375 */

/* the reference's 3x3 SVD (src/svd3.h:354, a non-inline function compiled into kernel.o) */
void svd(float a11, float a12, float a13, float a21, float a22, float a23, float a31, float a32, float a33,
         float &u11, float &u12, float &u13, float &u21, float &u22, float &u23, float &u31, float &u32, float &u33,
         float &s11, float &s12, float &s13, float &s21, float &s22, float &s23, float &s31, float &s32, float &s33,
         float &v11, float &v12, float &v13, float &v21, float &v22, float &v23, float &v31, float &v32, float &v33);

Lidar::Lidar(std::string) {}
Lidar::~Lidar() {}

#ifndef REF_SHIM_LIDAR_ONLY
static Patch make_patch(float sx, float sy, float rx, float ry)
{
    Patch p; p.scale = glm::vec3(sx, sy, 0.0f); p.resolution = glm::vec3(rx, ry, 1.0f);
    p.grid = NULL; p.uid = 0; return p;
}

extern "C" {

unsigned int ref_utilhash(unsigned int a) { return utilhash(a); }

int ref_sizeof_particle() { return (int)sizeof(Particle); }
int ref_sizeof_kdnode() { return (int)sizeof(KDTree::Node); }

void ref_clean_lidar_scan(int n, float scan, float theta, float *out_xy)
{
    glm::vec2 v; CleanLidarScan(n, scan, theta, v); out_xy[0] = v.x; out_xy[1] = v.y;
}

void ref_trace_ray(int sx, int sy, int ex, int ey, int w, int h, unsigned char *out)
{
    traceRay(glm::ivec2(sx, sy), glm::ivec2(ex, ey), glm::ivec2(w, h), (bool *)out);
}

/* EvaluateParticle for n particles given as SoA */
void ref_evaluate_particles(signed char *grid, int map_w, int map_h, float sx, float sy, float rx,
                            float ry, const float *x, const float *y, const float *th, int n,
                            float *scan, int *fit)
{
    Patch p = make_patch(sx, sy, rx, ry);
    for (int i = 0; i < n; i++) {
        Particle q; q.pos = glm::vec3(x[i], y[i], th[i]); q.w = 1.0f; q.cluster = 0; q.map = NULL;
        fit[i] = EvaluateParticle((MAP_TYPE *)grid, glm::ivec2(map_w, map_h), p, q, glm::vec3(0.0f), scan);
    }
}

/* ParticleAddNoise on particles idx0..idx0+n-1 */
void ref_add_noise(float *x, float *y, float *th, int n, int frame, int idx0)
{
    for (int i = 0; i < n; i++) {
        Particle q; q.pos = glm::vec3(x[i], y[i], th[i]); q.w = 1.0f; q.cluster = 0; q.map = NULL;
        ParticleAddNoise(q, frame, idx0 + i);
        x[i] = q.pos.x; y[i] = q.pos.y; th[i] = q.pos.z;
    }
}

/* Scene parser on the reference's own settings file: returns maps[0] scale/resolution */
int ref_scene_map(const char *path, float *out6)
{
    Scene *s = new Scene(std::string(path));
    if (s->maps.empty()) return -1;
    out6[0] = s->maps[0].scale.x; out6[1] = s->maps[0].scale.y; out6[2] = s->maps[0].scale.z;
    out6[3] = s->maps[0].resolution.x; out6[4] = s->maps[0].resolution.y; out6[5] = s->maps[0].resolution.z;
    return 0;
}

/* kd-tree host code (src/kdtree.cpp).  nodes: n x 8 int32/float32 words {axis,left,right,parent,x,y,z,w} */
void ref_kd_create(const float *pts4, int n, void *nodes)
{
    std::vector<glm::vec4> v(n);
    for (int i = 0; i < n; i++) v[i] = glm::vec4(pts4[4*i], pts4[4*i+1], pts4[4*i+2], pts4[4*i+3]);
    KDTree::Create(v, (KDTree::Node *)nodes);
}
void ref_kd_insert(const float *pt4, void *nodes, int size)
{
    KDTree::InsertNode(glm::vec4(pt4[0], pt4[1], pt4[2], pt4[3]), (KDTree::Node *)nodes, size);
}
void ref_kd_balance(void *nodes, int size) { KDTree::Balance((KDTree::Node *)nodes, size); }

/* Rotation the reference derives from the ICP cross-covariance (kernel.cu:1056-1079): feeds W to the
 * reference's svd() with the reference's argument order and returns R = U V^T (glm column-major, 9
 * floats) so the oracle's closed form can be compared against it. */
void ref_icp_rotation(const float *w9, float *r9)
{
    glm::mat3 W, U, S, V;
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) W[c][r] = w9[3 * c + r];
    svd(W[0][0], W[0][1], W[0][2], W[1][0], W[1][1], W[1][2], W[2][0], W[2][1], W[2][2],
        U[0][0], U[0][1], U[0][2], U[1][0], U[1][1], U[1][2], U[2][0], U[2][1], U[2][2],
        S[0][0], S[0][1], S[0][2], S[1][0], S[1][1], S[1][2], S[2][0], S[2][1], S[2][2],
        V[0][0], V[0][1], V[0][2], V[1][0], V[1][1], V[1][2], V[2][0], V[2][1], V[2][2]);
    glm::mat3 Um(glm::vec3(U[0][0], U[1][0], U[2][0]), glm::vec3(U[0][1], U[1][1], U[2][1]), glm::vec3(U[0][2], U[1][2], U[2][2]));
    glm::mat3 Vt(glm::vec3(V[0][0], V[0][1], V[0][2]), glm::vec3(V[1][0], V[1][1], V[1][2]), glm::vec3(V[2][0], V[2][1], V[2][2]));
    glm::mat3 R = Um * Vt;
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) r9[3 * c + r] = R[c][r];
}

}
#endif
