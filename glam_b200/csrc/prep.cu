// Parameter-space preparation for the triplet layers (tiny, one CTA each).
//
// The attention logit of TripletMessage is linear in the projected features (src_1gp/layer.py:48-49), so the
// per-node terms s_i = <a_i, xp_h>, s_j = <a_j, xp_h> are produced by the node projection itself: the 2H
// columns weight_node_h * a_{i|j},h are appended to weight_node, and the per-edge
// term becomes edge_attr @ (weight_edge_h * a_e,h).  Forward builds those derived weights, backward chains
// their gradients back to the reference's parameters.
#include "common.cuh"

namespace glam {

// att layout: full layer  [H][3C] = a_i | a_e | a_j ;  light layer [2C+De] = a_i | a_e(De) | a_j  (H == 1)
__global__ void triplet_prep_fwd_kernel(const float* __restrict__ wn, const float* __restrict__ we,
                                        const float* __restrict__ att, int C, int H, int De, int light, int ldxp,
                                        float* __restrict__ w_ext, float* __restrict__ att_edge) {
    const int HC = H * C;
    const int att_ld = light ? 2 * C + De : 3 * C;
    const int aj_off = light ? C + De : 2 * C;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < C * ldxp; idx += gridDim.x * blockDim.x) {
        const int k = idx / ldxp, n = idx - k * ldxp;
        float v = 0.f;
        if (n < HC) v = wn[k * HC + n];
        else if (n < HC + 2 * H) {
            const int which = (n - HC) / H, h = (n - HC) % H;
            const float* a = att + h * att_ld + (which == 0 ? 0 : aj_off);
            for (int c = 0; c < C; ++c) v = fmaf(wn[k * HC + h * C + c], a[c], v);
        }
        w_ext[idx] = v;
    }
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < De * H; idx += gridDim.x * blockDim.x) {
        const int d = idx / H, h = idx - d * H;
        float v = 0.f;
        if (light) v = att[C + d];
        else
            for (int c = 0; c < C; ++c) v = fmaf(we[d * HC + h * C + c], att[h * att_ld + C + c], v);
        att_edge[idx] = v;
    }
}

__global__ void triplet_prep_bwd_kernel(const float* __restrict__ wn, const float* __restrict__ we,
                                        const float* __restrict__ att, const float* __restrict__ g_w_ext,
                                        const float* __restrict__ g_att_edge, const float* __restrict__ g_we_direct,
                                        int C, int H, int De, int light, int ldxp, float* __restrict__ g_wn,
                                        float* __restrict__ g_we, float* __restrict__ g_att) {
    const int HC = H * C;
    const int att_ld = light ? 2 * C + De : 3 * C;
    const int aj_off = light ? C + De : 2 * C;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < C * HC; idx += gridDim.x * blockDim.x) {
        const int k = idx / HC, n = idx - k * HC, h = n / C, c = n - h * C;
        float v = g_w_ext[k * ldxp + n];
        v = fmaf(g_w_ext[k * ldxp + HC + h], att[h * att_ld + c], v);
        v = fmaf(g_w_ext[k * ldxp + HC + H + h], att[h * att_ld + aj_off + c], v);
        g_wn[idx] = v;
    }
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < H * C; idx += gridDim.x * blockDim.x) {
        const int h = idx / C, c = idx - h * C;
        float gi = 0.f, gj = 0.f;
        for (int k = 0; k < C; ++k) {
            const float w = wn[k * HC + h * C + c];
            gi = fmaf(w, g_w_ext[k * ldxp + HC + h], gi);
            gj = fmaf(w, g_w_ext[k * ldxp + HC + H + h], gj);
        }
        g_att[h * att_ld + c] = gi;
        g_att[h * att_ld + aj_off + c] = gj;
        if (!light) {
            float ge = 0.f;
            for (int d = 0; d < De; ++d) ge = fmaf(we[d * HC + h * C + c], g_att_edge[d * H + h], ge);
            g_att[h * att_ld + C + c] = ge;
        }
    }
    if (light) {
        for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < De; d += gridDim.x * blockDim.x) g_att[C + d] = g_att_edge[d];
    } else {
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < De * HC; idx += gridDim.x * blockDim.x) {
            const int d = idx / HC, n = idx - d * HC, h = n / C, c = n - h * C;
            g_we[idx] = fmaf(g_att_edge[d * H + h], att[h * att_ld + C + c], g_we_direct ? g_we_direct[idx] : 0.f);
        }
    }
}

}  // namespace glam

using namespace glam;

extern "C" int glam_triplet_prep_fwd(const float* weight_node, const float* weight_edge, const float* att, int C, int H,
                                     int De, int light, int ldxp, float* w_ext, float* att_edge, void* stream_) {
    GLAM_REQUIRE(C > 0 && H > 0 && De > 0 && ldxp >= H * C + 2 * H, "glam_triplet_prep_fwd: bad shape");
    GLAM_REQUIRE(!light || H == 1, "glam_triplet_prep_fwd: the Light layer is single-head");
    GLAM_REQUIRE(weight_node && att && w_ext && att_edge && (light || weight_edge), "glam_triplet_prep_fwd: null pointer");
    triplet_prep_fwd_kernel<<<(unsigned)((C * ldxp + 127) / 128), 128, 0, (cudaStream_t)stream_>>>(weight_node, weight_edge, att, C, H, De, light, ldxp, w_ext, att_edge);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_triplet_prep_bwd(const float* weight_node, const float* weight_edge, const float* att,
                                     const float* g_w_ext, const float* g_att_edge, const float* g_w_edge_direct, int C,
                                     int H, int De, int light, int ldxp, float* g_weight_node, float* g_weight_edge,
                                     float* g_att, void* stream_) {
    GLAM_REQUIRE(C > 0 && H > 0 && De > 0 && ldxp >= H * C + 2 * H, "glam_triplet_prep_bwd: bad shape");
    GLAM_REQUIRE(weight_node && att && g_w_ext && g_att_edge && g_weight_node && g_att, "glam_triplet_prep_bwd: null pointer");
    GLAM_REQUIRE(light || (weight_edge && g_weight_edge), "glam_triplet_prep_bwd: null edge pointers");
    triplet_prep_bwd_kernel<<<(unsigned)((C * H * C + 127) / 128), 128, 0, (cudaStream_t)stream_>>>(weight_node, weight_edge, att, g_w_ext, g_att_edge, g_w_edge_direct, C,
                                                                   H, De, light, ldxp, g_weight_node, g_weight_edge, g_att);
    GLAM_CHECK_LAUNCH();
    return 0;
}
