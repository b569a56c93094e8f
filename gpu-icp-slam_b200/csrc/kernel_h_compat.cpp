// kernel_h_compat.cpp -- the reference's own C++ entry points (src/kernel.h:14-24), implemented on
// top of the C ABI (include/pfslam.h), so that the reference's src/main.cpp links against this
// library instead of its kernel.cu.  Compiled against the reference's headers (kernel.h, scene.h,
// lidar.h, kdtree.hpp, sceneStructs.h) for the exact Scene / Lidar / Particle / KDTree::Node
// layouts; produces the same mangled symbols:
//   _Z18particleFilterInitP5Scene  _Z18particleFilterFreev  _Z14particleFilterP6uchar4iP5Lidar
//   _Z7drawMapP6uchar4  _Z9getPCDataPP8ParticlePPcPPN6KDTree4NodeEPiS8_RN3glm5tvec3IfLNS9_9precisionE0EEE
//   _Z20particleFilterInitPCv  _Z20particleFilterFreePCv
//
// Differences from the reference, all at the boundary: the particle count is the run-time value
// PFSLAM_PARTICLE_COUNT (environment, default 1000 = kernel.cu:30) instead of a #define; the 2D
// occupancy-grid step runs unless PFSLAM_PATH=kd selects the kd-tree point-cloud step -- the one the
// reference's HEAD calls (kernel.cu:1714-1745); drawMap is a no-op (rendering is out of scope,
// SURVEY section 2).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernel.h"
#include "../../include/pfslam.h"

static pfslam_engine *g_engine = NULL;
static Scene *hst_scene = NULL;
static std::vector<Particle> particles;          // getPCData lends these, like kernel.cu:61
static std::vector<MAP_TYPE> occupancyGrid;
static glm::vec3 robotPos(0.0f);
static int particle_count = 1000;
static bool kd_path = false;
static std::vector<unsigned char> kd;            // getPCData lends these too (kernel.cu:78, :810): raw KDTree::Node records

static void die(const char *what)
{
    // same contract as checkCUDAError (kernel.h:42-59): report and exit
    fprintf(stderr, "pfslam error: %s: %s\n", what, pfslam_last_error());
    exit(EXIT_FAILURE);
}

void particleFilterInitPC() {}
void particleFilterFreePC() {}

void particleFilterInit(Scene *scene)
{
    hst_scene = scene;
    const char *env = getenv("PFSLAM_PARTICLE_COUNT");
    if (env && atoi(env) > 0) particle_count = atoi(env);
    pfslam_config cfg;
    pfslam_default_config(&cfg);
    cfg.n_particles = cfg.n_particles_global = particle_count;
    const Patch &m = scene->maps[0];                 // kernel.cu:119
    cfg.map_scale_x = m.scale.x; cfg.map_scale_y = m.scale.y;
    cfg.map_res_x = m.resolution.x; cfg.map_res_y = m.resolution.y;
    const char *path = getenv("PFSLAM_PATH");
    kd_path = path && strcmp(path, "kd") == 0;
    if (path && !kd_path && strcmp(path, "grid2d") != 0) {
        fprintf(stderr, "pfslam error: PFSLAM_PATH must be grid2d or kd, not '%s'\n", path);
        exit(EXIT_FAILURE);
    }
    cfg.path = kd_path ? PFSLAM_PATH_KD : PFSLAM_PATH_GRID2D;
    const char *dev = getenv("PFSLAM_DEVICE");
    if (dev) cfg.device = atoi(dev);
    if (pfslam_create(&cfg, &g_engine) != PFSLAM_OK) die("particleFilterInit");
    particles.assign(particle_count, Particle());
    int w = 0, h = 0;
    pfslam_get_map_dim(g_engine, &w, &h);
    occupancyGrid.assign((size_t)w * h, (MAP_TYPE)-100);
    robotPos = glm::vec3(0.0f);
    particleFilterInitPC();
}

void particleFilterFree()
{
    pfslam_destroy(g_engine);                        // NULL-safe: called before Init (main.cpp:194)
    g_engine = NULL;
    particleFilterFreePC();
}

void particleFilter(uchar4 *pbo, int frame, Lidar *lidar)
{
    (void)pbo;                                       // unused at the reference's HEAD too
    pfslam_frame_result r;
    if (pfslam_step(g_engine, lidar->scans[frame].data(), frame, &r) != PFSLAM_OK) die("particleFilter");
    robotPos = glm::vec3(r.pose[0], r.pose[1], r.pose[2]);
}

void drawMap(uchar4 *pbo) { (void)pbo; }

void getPCData(Particle **ptrParticles, MAP_TYPE **ptrMap, KDTree::Node **ptrKD, int *nParticles, int *nKD,
               glm::vec3 &pos)
{
    std::vector<float> x(particle_count), y(particle_count), t(particle_count), w(particle_count);
    if (pfslam_get_particles(g_engine, x.data(), y.data(), t.data(), w.data()) != PFSLAM_OK) die("getPCData");
    for (int i = 0; i < particle_count; i++) {
        particles[i].pos = glm::vec3(x[i], y[i], t[i]);
        particles[i].w = w[i];
        particles[i].cluster = 0;
        particles[i].map = NULL;
    }
    if (pfslam_get_grid(g_engine, (int8_t *)occupancyGrid.data()) != PFSLAM_OK) die("getPCData");
    *ptrParticles = particles.data();
    *ptrMap = occupancyGrid.data();
    *nParticles = particle_count;
    *ptrKD = NULL;                                   // 2D path: no kd nodes
    *nKD = 0;
    if (kd_path) {                                   // kernel.cu:810-811: the host copy of the tree and its size
        int32_t n = 0;
        static_assert(sizeof(KDTree::Node) == 32, "KDTree::Node layout");
        if (pfslam_get_kd(g_engine, NULL, 0, &n) != PFSLAM_OK) die("getPCData");
        if ((size_t)n * sizeof(KDTree::Node) > kd.size()) kd.resize(((size_t)n + 4096) * sizeof(KDTree::Node));
        if (n > 0 && pfslam_get_kd(g_engine, kd.data(), n, &n) != PFSLAM_OK) die("getPCData");
        *ptrKD = reinterpret_cast<KDTree::Node *>(kd.data());
        *nKD = n;
    }
    pos = robotPos;
}
