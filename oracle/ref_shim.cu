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

Lidar::Lidar(std::string) {}
Lidar::~Lidar() {}

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

}
