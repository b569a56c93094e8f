// pfslam_run.cpp -- headless replacement for the loop of the reference's src/main.cpp.
//
// Calls the reference's own entry points (src/kernel.h:14-24) in the order of runCuda()
// (main.cpp:175-237): particleFilterFree(); particleFilterInit(scene); then for frame = 1..F-1
// particleFilter(pbo, frame, lidar); getPCData(...) at the end.  Linked against
// libpfslam_kernelh.so + libpfslam.so instead of the reference's kernel.cu, and compiled against the
// reference's headers, it is the drop-in demonstration: the only things replaced are the GL/PCL shell
// and the MathWorks loader (scans come from the packed .scans.u16 format).
//
//   pfslam_run <map_settings.txt> <file.scans.u16> [max_frames] [trajectory.csv]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernel.h"

Lidar::Lidar(string filename)           // src/lidar.cpp needs libmat; read the packed format instead
{
    FILE *f = fopen(filename.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", filename.c_str()); throw 1; }
    char magic[8]; uint32_t nf = 0, nb = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "PFSCAN1", 8) != 0 || fread(&nf, 4, 1, f) != 1 || fread(&nb, 4, 1, f) != 1) {
        fprintf(stderr, "%s is not a PFSCAN1 file\n", filename.c_str()); throw 1;
    }
    std::vector<uint16_t> row(nb);
    for (uint32_t i = 0; i < nf; i++) {
        if (fread(row.data(), 2, nb, f) != nb) break;
        std::vector<float> s(nb);
        for (uint32_t j = 0; j < nb; j++) s[j] = row[j] == 65535 ? 4294967.0f : (float)((double)row[j] / 1000.0);
        scans.push_back(s);
    }
    fclose(f);
}
Lidar::~Lidar() {}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <map_settings.txt> <file.scans.u16> [max_frames] [trajectory.csv]\n", argv[0]); return 2; }
    Scene *scene = new Scene(argv[1]);
    Lidar *lidar = new Lidar(argv[2]);
    int last = (int)lidar->scans.size() - 1;
    if (argc > 3 && atoi(argv[3]) > 0 && atoi(argv[3]) < last) last = atoi(argv[3]);
    FILE *csv = argc > 4 ? fopen(argv[4], "w") : NULL;

    particleFilterFree();                       // main.cpp:194
    particleFilterInit(scene);                  // main.cpp:195
    Particle *particles; MAP_TYPE *map; KDTree::Node *kd; int nParticles = 0, nKD = 0; glm::vec3 pos;
    for (int frame = 1; frame <= last; frame++) {   // main.cpp:199-206: scan 0 is never consumed
        particleFilter(NULL, frame, lidar);
        if (csv) {
            getPCData(&particles, &map, &kd, &nParticles, &nKD, pos);
            fprintf(csv, "%d,%.9g,%.9g,%.9g\n", frame, pos.x, pos.y, pos.z);
        }
    }
    getPCData(&particles, &map, &kd, &nParticles, &nKD, pos);
    long occupied = 0, seen = 0;
    for (long i = 0; i < 1600L * 1600L; i++) { if (map[i] != -100) seen++; if (map[i] > 0) occupied++; }
    double kd_w = 0.0;
    for (int i = 0; i < nKD; i++) kd_w += kd[i].value.w;
    printf("frames %d particles %d robotPos %.9g %.9g %.9g cells_seen %ld cells_occupied %ld kd_nodes %d kd_weight_sum %.0f\n", last, nParticles,
           pos.x, pos.y, pos.z, seen, occupied, nKD, kd_w);
    if (csv) fclose(csv);
    particleFilterFree();
    return 0;
}
