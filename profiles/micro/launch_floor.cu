// launch_floor.cu — what does an (almost) empty launch cost inside a CUDA-graph chain on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_floor launch_floor.cu && ./launch_floor
// Variants: grid shape, dynamic shared memory, kernel-parameter size, programmatic dependent launch, a zero-filled
// shared-memory prologue, and one 4 KB TMA-less global load + store per CTA.  Prints us per launch.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <vector>

struct Small { int* out; int n; int zero_bytes; int touch; };
struct Big { Small s; char pad[1500]; };

template <typename P>
__global__ void __launch_bounds__(1024) k(const __grid_constant__ P prm) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Small& s = *reinterpret_cast<const Small*>(&prm);
    if (s.zero_bytes) {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* q = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < (s.zero_bytes >> 4); i += blockDim.x) q[i] = z;
        __syncthreads();
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (s.touch) {                                   // one coalesced read + write per thread
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        s.out[i] = s.out[i] + 1;
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename P>
static float run(const P& prm, int grid, int block, int smem, bool pdl, int n_nodes = 56, int replays = 400) {
    cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < n_nodes; i++) {
        cudaLaunchConfig_t lc; memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3(grid); lc.blockDim = dim3(block); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&lc, k<P>, prm);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int r = 0; r < replays; r++) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("  (error: %s)\n", cudaGetErrorString(err));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    return ms * 1e3f / (replays * n_nodes);
}

int main() {
    int* buf; cudaMalloc(&buf, 1 << 26); cudaMemset(buf, 0, 1 << 26);
    Small s{buf, 0, 0, 0};
    Big b; memset(&b, 0, sizeof(b)); b.s = s;
    printf("%-70s %8s %8s\n", "variant (us per launch, 56-node graph chain)", "no PDL", "PDL");
    struct V { const char* name; int grid, block, smem, zero, touch, big; };
    std::vector<V> vs = {
        {"1 CTA x 32 thr, no smem", 1, 32, 0, 0, 0, 0},
        {"148 CTAs x 64 thr, no smem", 148, 64, 0, 0, 0, 0},
        {"2048 CTAs x 64 thr, no smem", 2048, 64, 0, 0, 0, 0},
        {"2048 CTAs x 64 thr, 13 KB smem", 2048, 64, 13056, 0, 0, 0},
        {"2048 CTAs x 64 thr, 13 KB smem, 1.5 KB params", 2048, 64, 13056, 0, 0, 1},
        {"2048 CTAs x 64 thr, 13 KB smem, 1.5 KB params, zero 8 KB", 2048, 64, 13056, 8064, 0, 1},
        {"2048 CTAs x 64 thr, 13 KB smem, 1.5 KB params, zero 8 KB, touch", 2048, 64, 13056, 8064, 1, 1},
        {"1024 CTAs x 128 thr, 26 KB smem, 1.5 KB params, zero 16 KB", 1024, 128, 26112, 16128, 0, 1},
        {"512 CTAs x 256 thr, 52 KB smem, 1.5 KB params, zero 32 KB", 512, 256, 52224, 32256, 0, 1},
        {"293 CTAs x 448 thr, 92 KB smem, 1.5 KB params, zero 56 KB", 293, 448, 91392 + 256, 56448, 0, 1},
        {"293 CTAs x 448 thr, 92 KB smem, small params, zero 56 KB", 293, 448, 91392 + 256, 56448, 0, 0},
        {"293 CTAs x 448 thr, 92 KB smem, 1.5 KB params, no zero", 293, 448, 91392 + 256, 0, 0, 1},
        {"148 CTAs x 896 thr, 183 KB smem, 1.5 KB params, zero 113 KB", 148, 896, 182784 + 256, 112896, 0, 1},
        {"148 CTAs x 896 thr, 183 KB smem, 1.5 KB params, no zero", 148, 896, 182784 + 256, 0, 0, 1},
        {"296 CTAs x 448 thr, no smem, small params", 296, 448, 0, 0, 0, 0},
    };
    for (auto& v : vs) {
        float t[2];
        for (int pdl = 0; pdl < 2; pdl++) {
            if (v.big) { Big p = b; p.s.zero_bytes = v.zero; p.s.touch = v.touch; t[pdl] = run(p, v.grid, v.block, v.smem, pdl); }
            else { Small p = s; p.zero_bytes = v.zero; p.touch = v.touch; t[pdl] = run(p, v.grid, v.block, v.smem, pdl); }
        }
        printf("%-70s %8.2f %8.2f\n", v.name, t[0], t[1]);
    }
    return 0;
}
