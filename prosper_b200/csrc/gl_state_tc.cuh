// Tensor-core state kernel of the binary Gaussian-linear model (gl_state_tc.cu): constant tables and launcher.
#pragma once
#include <vector>

#include "gl_kernel.cuh"

namespace pet {

constexpr int TC_NC = 64;                          // states per chunk
constexpr int TC_NOUT = 80;                        // reverse-product outputs (features + spare), multiple of 16
constexpr int TC_KF = 96;                          // padded feature count
constexpr int TC_BFWD_BYTES = 2 * TC_KF * TC_NC;   // forward membership operand of one chunk (weights 1 and 128)
constexpr int TC_BREV_BYTES = 2 * TC_NC * TC_NOUT; // reverse membership operand of one chunk
constexpr int TC_MAX_CHUNKS = 96;

struct GLTc {                                      // kernel parameter
    int n_chunks, n_feat, max_nfeat;
    const uint8_t *bfwd, *brev;                    // device images, n_chunks x TC_BFWD_BYTES / TC_BREV_BYTES
    uint8_t chunk_cnt[TC_MAX_CHUNKS];              // valid states of a chunk (the rest is padding)
    uint8_t feat[2 * TC_KF];                       // feature f -> candidate positions (j, k); j == k: linear term
};

struct GLTcHost {
    GLTc dev;
    std::vector<uint8_t> bfwd, brev;
};

bool gl_tc_supported(const GLStatic &st, int gamma, bool binary);
int gl_tc_build_tables(const GLStatic &st, int gamma, const std::vector<double> &matrix, GLTcHost &out);
int launch_gl_state_tc(const GLArgs &a, const GLTc &t, int sm_count, cudaStream_t stream);

}  // namespace pet
