
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

}
