
/* ===== appended by oracle/Makefile (T3): test-only access to the reference's file-static device
 * state, so that the reference's OWN CUDA kernels can be run on the GPU box on chosen inputs and
 * their outputs read back.  Everything above this line is the unmodified reference kernel.cu,
 * concatenated at build time from /root/reference/src (never copied into the repository). ===== */
extern "C" {

int t3_particle_count() { return PARTICLE_COUNT; }

/* particleFilterInit needs a Scene; build one from the reference's own settings file */
int t3_init(const char *scene_path)
{
    Scene *s = new Scene(std::string(scene_path));
    particleFilterFree();
    particleFilterInit(s);
    return (int)cudaGetLastError();
}

void t3_set_particles(const float *x, const float *y, const float *th, const float *w)
{
    for (int i = 0; i < PARTICLE_COUNT; i++) { particles[i].pos = glm::vec3(x[i], y[i], th[i]); particles[i].w = w[i]; }
    cudaMemcpy(dev_particles, particles, PARTICLE_COUNT * sizeof(Particle), cudaMemcpyHostToDevice);
}
void t3_get_particles(float *x, float *y, float *th, float *w)
{
    cudaMemcpy(particles, dev_particles, PARTICLE_COUNT * sizeof(Particle), cudaMemcpyDeviceToHost);
    for (int i = 0; i < PARTICLE_COUNT; i++) { x[i] = particles[i].pos.x; y[i] = particles[i].pos.y; th[i] = particles[i].pos.z; w[i] = particles[i].w; }
}
void t3_set_grid(const signed char *g) { cudaMemcpy(dev_occupancyGrid, g, map_dim.x * map_dim.y, cudaMemcpyHostToDevice); }
void t3_get_grid(signed char *g) { cudaMemcpy(g, dev_occupancyGrid, map_dim.x * map_dim.y, cudaMemcpyDeviceToHost); }
void t3_set_robot(float x, float y, float th) { robotPos = glm::vec3(x, y, th); }
void t3_get_robot(float *p) { p[0] = robotPos.x; p[1] = robotPos.y; p[2] = robotPos.z; }
void t3_get_fit(int *fit) { cudaMemcpy(fit, dev_fit, PARTICLE_COUNT * sizeof(int), cudaMemcpyDeviceToHost); }

/* the reference's own step functions, GPU branches (GPU_* == 1) */
void t3_motion(int frame) { PFMotionUpdate(frame); }
void t3_measure(const float *scan, float *pose)
{
    std::vector<float> v(scan, scan + LIDAR_SIZE);
    glm::vec3 p = PFMeasurementUpdate(v);
    pose[0] = p.x; pose[1] = p.y; pose[2] = p.z;
}
void t3_update_map(const float *scan) { std::vector<float> v(scan, scan + LIDAR_SIZE); PFUpdateMap(v); cudaDeviceSynchronize(); }
void t3_resample(int frame) { PFResample(frame); cudaDeviceSynchronize(); }
int t3_last_error() { return (int)cudaGetLastError(); }


/* ---- kd-tree point-cloud path (kernel.cu:816-1540, :1702-1768) ---- */
void t3_set_alloc_fill(int byte);                  /* oracle/ref_t3_alloc.cu */
void t3_set_alloc_fill_sized(int byte, size_t size, int byte_special);
void t3_reset_kd() { kdSize = 0; }                 /* particleFilterInit does not */
int t3_kd_size() { return kdSize; }
void t3_kd_set(const void *nodes, int n)
{
    memcpy(kd, nodes, (size_t)n * sizeof(KDTree::Node));
    kdSize = n;
    cudaMemcpy(dev_kd, kd, (size_t)n * sizeof(KDTree::Node), cudaMemcpyHostToDevice);
}
/* the device copy carries the weights (the host copy is refreshed only when a node is inserted) */
void t3_kd_get(void *nodes) { cudaMemcpy(nodes, dev_kd, (size_t)kdSize * sizeof(KDTree::Node), cudaMemcpyDeviceToHost); }
void t3_get_fitf(float *fit) { cudaMemcpy(fit, dev_fitf, PARTICLE_COUNT * sizeof(float), cudaMemcpyDeviceToHost); }

/* findCorrespondenceIndexKD (kernel.cu:924) on n query points (x, y, z, w) */
void t3_kd_nn(const float *q4, int n, int *idx)
{
    glm::vec4 *dq = NULL; int *di = NULL;
    cudaMalloc((void **)&dq, n * sizeof(glm::vec4));
    cudaMalloc((void **)&di, n * sizeof(int));
    cudaMemcpy(dq, q4, n * sizeof(glm::vec4), cudaMemcpyHostToDevice);
    findCorrespondenceIndexKD<<<(n + BLOCK_SIZE - 1) / BLOCK_SIZE, BLOCK_SIZE>>>(n, di, dq, dev_kd);
    cudaDeviceSynchronize();
    cudaMemcpy(idx, di, n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dq); cudaFree(di);
}
/* robotPos = PFMeasurementUpdateKD(scan), as particleFilter() does (kernel.cu:1736) */
void t3_measure_kd(const float *scan, float *pose)
{
    std::vector<float> v(scan, scan + LIDAR_SIZE);
    robotPos = PFMeasurementUpdateKD(v);
    pose[0] = robotPos.x; pose[1] = robotPos.y; pose[2] = robotPos.z;
}
void t3_icp(const float *start, const float *scan, float *out)
{
    std::vector<float> v(scan, scan + LIDAR_SIZE);
    glm::vec3 p = transformPointICP(glm::vec3(start[0], start[1], start[2]), v);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void t3_update_map_kd(const float *scan) { std::vector<float> v(scan, scan + LIDAR_SIZE); PFUpdateMapKD(v); cudaDeviceSynchronize(); }
/* the reference's per-frame driver itself (kernel.cu:1702-1768, the kd step at HEAD) */
void t3_particle_filter(const float *scan, int frame)
{
    static Lidar *L = new Lidar(std::string(""));
    if ((int)L->scans.size() <= frame) L->scans.resize(frame + 1);
    L->scans[frame].assign(scan, scan + LIDAR_SIZE);
    particleFilter(NULL, frame, L);
    cudaDeviceSynchronize();
}


/* ---- timing accessors for bench.py's reference_gpu leg (the "beat this kernel" bar) ---- */
static double t3_now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
/* one 2D occupancy-grid step in the README.md:41-50 order, timed the reference's own way (wall clock
 * around blocking calls, kernel.cu:1727-1748); phase_ms[4] = motion, measurement, map, resample */
double t3_step2d_ms(const float *scan, int frame, double *phase_ms)
{
    std::vector<float> v(scan, scan + LIDAR_SIZE);
    cudaDeviceSynchronize();
    const double t0 = t3_now_ms();
    PFMotionUpdate(frame);
    const double t1 = t3_now_ms();
    robotPos = PFMeasurementUpdate(v);
    const double t2 = t3_now_ms();
    PFUpdateMap(v);
    cudaDeviceSynchronize();
    const double t3 = t3_now_ms();
    PFResample(frame);
    cudaDeviceSynchronize();
    const double t4 = t3_now_ms();
    if (phase_ms) { phase_ms[0] = t1 - t0; phase_ms[1] = t2 - t1; phase_ms[2] = t3 - t2; phase_ms[3] = t4 - t3; }
    return t4 - t0;
}
/* the kd step at HEAD through the reference's own driver, wall clock */
double t3_step_kd_ms(const float *scan, int frame)
{
    cudaDeviceSynchronize();
    const double t0 = t3_now_ms();
    t3_particle_filter(scan, frame);
    return t3_now_ms() - t0;
}
/* kernEvaluateParticles (kd != 0: kernEvaluateParticlesKD) alone, CUDA events, mean of reps launches */
float t3_time_evaluate_ms(const float *scan, int reps, int kd_path)
{
    const int blockSize1d = 128;
    const dim3 blocksPerGrid1d((PARTICLE_COUNT + blockSize1d - 1) / blockSize1d);
    cudaMemcpy(dev_lidar, scan, LIDAR_SIZE * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(dev_particles, particles, PARTICLE_COUNT * sizeof(Particle), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float tot = 0.0f;
    for (int r = 0; r <= reps; r++) {             /* launch 0 is a warm-up */
        cudaEventRecord(e0);
        if (kd_path)
            kernEvaluateParticlesKD<<<blocksPerGrid1d, blockSize1d>>>(dev_occupancyGrid, map_dim, map_params, dev_particles, robotPos, dev_lidar, dev_fitf, dev_kd, kdSize);
        else
            kernEvaluateParticles<<<blocksPerGrid1d, blockSize1d>>>(dev_occupancyGrid, map_dim, map_params, dev_particles, robotPos, dev_lidar, dev_fit);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r) tot += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return tot / (float)reps;
}

}
