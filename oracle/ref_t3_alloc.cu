/*
 * ref_t3_alloc.cu -- TEST INFRASTRUCTURE (T3): defines the memory the reference reads without having
 * written it, so that its kd kernels can be compared with the oracle on the GPU box.
 *
 * oracle/Makefile links libref_t3.so with  --wrap=cudaMalloc --wrap=cudaFree : every cudaMalloc the
 * reference's unmodified kernel.cu (and Thrust inside it) makes comes through here.  Nothing of the
 * reference's source is changed; only the CONTENT of fresh device memory is made deterministic:
 *
 *   * every block gets a 256-byte front pad holding eight "stop" nodes {axis 3, left -1, right -1,
 *     parent -1, value 0}.  The reference's NN walk reads tree[tree[best].parent] with parent == -1 when
 *     the root is the best node (kernel.cu:911, :961, :1176, :1268 -- SURVEY Q9), i.e. the 32 bytes in
 *     front of dev_kd.  With the pad that record has no valid axis (getHyperplaneDist returns 0 and leaves
 *     `branch` alone) and both links are -1, so the walk ends there -- the oracle's definition of Q9.
 *   * the body is filled with the byte chosen by t3_set_alloc_fill (default: left as cudaMalloc returns
 *     it).  0xFF makes the uninitialised tail of dev_free (kernel.cu:1475 copies wallPC.size() points
 *     into a freePC.size() buffer -- Q10) NaN points, which update nothing: the oracle's "only the first
 *     |wallPC| free points are applied".  0x00 makes the never-written ICP targets of out-of-range beams
 *     (kernel.cu:984-990 -- Q11) (0,0,0,0): the oracle's definition.
 */
#include <cstring>
#include <set>
#include <cuda_runtime.h>

extern "C" {

cudaError_t __real_cudaMalloc(void **p, size_t n);
cudaError_t __real_cudaFree(void *p);

static int g_fill = -1, g_fill_special = -1;
static size_t g_special_size = 0;
static std::set<void *> *g_live = nullptr;
enum { kPad = 256 };

void t3_set_alloc_fill(int byte) { g_fill = byte; g_special_size = 0; }
/* blocks of exactly `size` bytes get `byte_special` instead (free-running kd runs need NaN for the dev_free
 * tail and zeros for the LIDAR_SIZE * sizeof(vec4) ICP target buffer in the same frame) */
void t3_set_alloc_fill_sized(int byte, size_t size, int byte_special) { g_fill = byte; g_special_size = size; g_fill_special = byte_special; }

cudaError_t __wrap_cudaMalloc(void **p, size_t n)
{
    if (!g_live) g_live = new std::set<void *>();
    void *base = nullptr;
    cudaError_t e = __real_cudaMalloc(&base, n + kPad);
    if (e != cudaSuccess) return e;
    int pad[kPad / 4];
    for (int i = 0; i < kPad / 32; i++) {
        int *q = pad + 8 * i;
        q[0] = 3; q[1] = -1; q[2] = -1; q[3] = -1; q[4] = q[5] = q[6] = q[7] = 0;
    }
    cudaMemcpy(base, pad, kPad, cudaMemcpyHostToDevice);
    const int fill = (g_special_size && n == g_special_size) ? g_fill_special : g_fill;
    if (fill >= 0 && n > 0) cudaMemset((char *)base + kPad, fill, n);
    *p = (char *)base + kPad;
    g_live->insert(*p);
    return cudaSuccess;
}

cudaError_t __wrap_cudaFree(void *p)
{
    if (p && g_live) {
        auto it = g_live->find(p);
        if (it != g_live->end()) { g_live->erase(it); return __real_cudaFree((char *)p - kPad); }
    }
    return __real_cudaFree(p);
}

}
