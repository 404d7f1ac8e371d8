// Recurrent cell gates: GRU node update of MessageBlock and the LSTM cell of Set2Set
// (see include/glam_b200.h (4)).  Pure streaming kernels: one thread per (row, channel).
#include "common.cuh"

namespace glam {

__global__ void gru_gates_fwd_kernel(float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h,
                                     const float* __restrict__ identity, int64_t N, int C, int act, float act_param,
                                     float* __restrict__ h_new, float* __restrict__ x_out) {
    const int64_t total = N * C;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / C;
        const int c = (int)(idx - n * C);
        const int64_t b = n * 3 * C + c;
        const float r = sigmoidf_(gi[b] + gh[b]);
        const float z = sigmoidf_(gi[b + C] + gh[b + C]);
        const float nn = tanhf(gi[b + 2 * C] + r * gh[b + 2 * C]);
        const float hv = h[idx];
        const float hn = (1.f - z) * nn + z * hv;
        gi[b] = r; gi[b + C] = z; gi[b + 2 * C] = nn;
        h_new[idx] = hn;
        const float s = identity ? hn + identity[idx] : hn;
        x_out[idx] = act_fwd(s, act, act_param);
    }
}

__global__ void gru_gates_bwd_kernel(const float* __restrict__ rzn, const float* __restrict__ gh_n, int64_t ldghn, const float* __restrict__ h,
                                     const float* __restrict__ x_out, const float* __restrict__ g_x_out,
                                     const float* __restrict__ g_h_carry, int64_t N, int C, int act, float act_param,
                                     float* __restrict__ g_gi, float* __restrict__ g_gh, float* __restrict__ g_h_prev,
                                     float* __restrict__ g_identity) {
    const int64_t total = N * C;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / C;
        const int c = (int)(idx - n * C);
        const int64_t b = n * 3 * C + c;
        const float r = rzn[b], z = rzn[b + C], nn = rzn[b + 2 * C], ghn = gh_n[n * ldghn + c], hv = h[idx];
        const float gs = g_x_out ? g_x_out[idx] * act_grad_from_out(x_out[idx], act, act_param) : 0.f;
        if (g_identity) g_identity[idx] = gs;
        const float ghp = gs + (g_h_carry ? g_h_carry[idx] : 0.f);
        const float g_n = ghp * (1.f - z);
        const float g_z = ghp * (hv - nn);
        const float g_npre = g_n * (1.f - nn * nn);
        const float g_zpre = g_z * z * (1.f - z);
        const float g_rpre = g_npre * ghn * r * (1.f - r);
        g_h_prev[idx] = ghp * z;
        g_gi[b] = g_rpre; g_gi[b + C] = g_zpre; g_gi[b + 2 * C] = g_npre;
        g_gh[b] = g_rpre; g_gh[b + C] = g_zpre; g_gh[b + 2 * C] = g_npre * r;
    }
}

__global__ void lstm_gates_fwd_kernel(float* __restrict__ gates, const float* __restrict__ c_prev, int64_t R, int C,
                                      float* __restrict__ c_new, float* __restrict__ h_new) {
    const int64_t total = R * C;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / C;
        const int c = (int)(idx - n * C);
        const int64_t b = n * 4 * C + c;
        const float i = sigmoidf_(gates[b]), f = sigmoidf_(gates[b + C]);
        const float g = tanhf(gates[b + 2 * C]), o = sigmoidf_(gates[b + 3 * C]);
        const float cn = f * c_prev[idx] + i * g;
        gates[b] = i; gates[b + C] = f; gates[b + 2 * C] = g; gates[b + 3 * C] = o;
        c_new[idx] = cn;
        h_new[idx] = o * tanhf(cn);
    }
}

__global__ void lstm_gates_bwd_kernel(const float* __restrict__ ga, const float* __restrict__ c_prev,
                                      const float* __restrict__ c_new, const float* __restrict__ g_h,
                                      const float* __restrict__ g_c_in, int64_t R, int C, float* __restrict__ g_gates,
                                      float* __restrict__ g_c_prev) {
    const int64_t total = R * C;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / C;
        const int c = (int)(idx - n * C);
        const int64_t b = n * 4 * C + c;
        const float i = ga[b], f = ga[b + C], g = ga[b + 2 * C], o = ga[b + 3 * C];
        const float tc = tanhf(c_new[idx]);
        const float gh = g_h ? g_h[idx] : 0.f;
        const float gc = (g_c_in ? g_c_in[idx] : 0.f) + gh * o * (1.f - tc * tc);
        g_gates[b] = gc * g * i * (1.f - i);
        g_gates[b + C] = gc * c_prev[idx] * f * (1.f - f);
        g_gates[b + 2 * C] = gc * i * (1.f - g * g);
        g_gates[b + 3 * C] = gh * tc * o * (1.f - o);
        g_c_prev[idx] = gc * f;
    }
}


// ---- 16-byte vectorised variants (channels % 4 == 0, 16-byte aligned rows): one thread = 4 channels of one node ----
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#define GLAM_F4_MAP(out, expr) { out.x = (expr(x)); out.y = (expr(y)); out.z = (expr(z)); out.w = (expr(w)); }

__global__ void __launch_bounds__(256)
gru_gates_fwd_vec_kernel(float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h,
                         const float* __restrict__ identity, int64_t N, int C, int act, float act_param,
                         float* __restrict__ h_new, float* __restrict__ x_out) {
    const int cq = C >> 2;
    const int64_t total = N * cq;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / cq;
        const int c = (int)(idx - n * cq) << 2;
        const int64_t b = n * 3 * C + c, o = n * C + c;
        const float4 ir = ld4(gi + b), iz = ld4(gi + b + C), in = ld4(gi + b + 2 * C);
        const float4 hr = ld4(gh + b), hz = ld4(gh + b + C), hn = ld4(gh + b + 2 * C);
        const float4 hv = ld4(h + o);
        float4 idv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (identity) idv = ld4(identity + o);
        float4 r, z, nn, hnw, xo;
#define R_(k) sigmoidf_(ir.k + hr.k)
        GLAM_F4_MAP(r, R_)
#define Z_(k) sigmoidf_(iz.k + hz.k)
        GLAM_F4_MAP(z, Z_)
#define N_(k) tanhf(in.k + r.k * hn.k)
        GLAM_F4_MAP(nn, N_)
#define H_(k) ((1.f - z.k) * nn.k + z.k * hv.k)
        GLAM_F4_MAP(hnw, H_)
#define X_(k) act_fwd(hnw.k + idv.k, act, act_param)
        GLAM_F4_MAP(xo, X_)
#undef R_
#undef Z_
#undef N_
#undef H_
#undef X_
        st4(gi + b, r); st4(gi + b + C, z); st4(gi + b + 2 * C, nn);
        st4(h_new + o, hnw);
        st4(x_out + o, xo);
    }
}

__global__ void __launch_bounds__(256)
gru_gates_bwd_vec_kernel(const float* __restrict__ rzn, const float* __restrict__ gh_n, int64_t ldghn, const float* __restrict__ h,
                         const float* __restrict__ x_out, const float* __restrict__ g_x_out,
                         const float* __restrict__ g_h_carry, int64_t N, int C, int act, float act_param,
                         float* __restrict__ g_gi, float* __restrict__ g_gh, float* __restrict__ g_h_prev,
                         float* __restrict__ g_identity) {
    const int cq = C >> 2;
    const int64_t total = N * cq;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = idx / cq;
        const int c = (int)(idx - n * cq) << 2;
        const int64_t b = n * 3 * C + c, o = n * C + c;
        const float4 r = ld4(rzn + b), z = ld4(rzn + b + C), nn = ld4(rzn + b + 2 * C), ghn = ld4(gh_n + n * ldghn + c), hv = ld4(h + o);
        float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g_x_out) {
            const float4 gx = ld4(g_x_out + o), xo = ld4(x_out + o);
#define G_(k) (gx.k * act_grad_from_out(xo.k, act, act_param))
            GLAM_F4_MAP(gs, G_)
#undef G_
        }
        if (g_identity) st4(g_identity + o, gs);
        float4 ghp = gs;
        if (g_h_carry) { const float4 cr = ld4(g_h_carry + o); ghp.x += cr.x; ghp.y += cr.y; ghp.z += cr.z; ghp.w += cr.w; }
        float4 gnp, gzp, grp, ghprev, gnr;
#define NP_(k) (ghp.k * (1.f - z.k) * (1.f - nn.k * nn.k))
        GLAM_F4_MAP(gnp, NP_)
#define ZP_(k) (ghp.k * (hv.k - nn.k) * z.k * (1.f - z.k))
        GLAM_F4_MAP(gzp, ZP_)
#define RP_(k) (gnp.k * ghn.k * r.k * (1.f - r.k))
        GLAM_F4_MAP(grp, RP_)
#define HP_(k) (ghp.k * z.k)
        GLAM_F4_MAP(ghprev, HP_)
#define NR_(k) (gnp.k * r.k)
        GLAM_F4_MAP(gnr, NR_)
#undef NP_
#undef ZP_
#undef RP_
#undef HP_
#undef NR_
        st4(g_h_prev + o, ghprev);
        st4(g_gi + b, grp); st4(g_gi + b + C, gzp); st4(g_gi + b + 2 * C, gnp);
        st4(g_gh + b, grp); st4(g_gh + b + C, gzp); st4(g_gh + b + 2 * C, gnr);
    }
}

static bool al16(const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; }

static int ew_grid(int64_t total) {
    int64_t g = (total + 255) / 256;
    int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace glam

using namespace glam;

extern "C" int glam_gru_gates_fwd(float* gi_rzn, const float* gh, const float* h, const float* identity, int64_t N, int C,
                                  int act, float act_param, float* h_new, float* x_out, void* stream_) {
    GLAM_REQUIRE(N >= 0 && C > 0 && act >= 0 && act <= 3, "glam_gru_gates_fwd: bad arguments");
    if (N == 0) return 0;
    GLAM_REQUIRE(gi_rzn && gh && h && h_new && x_out, "glam_gru_gates_fwd: null pointer");
    if ((C & 3) == 0 && al16(gi_rzn) && al16(gh) && al16(h) && al16(identity) && al16(h_new) && al16(x_out))
        gru_gates_fwd_vec_kernel<<<ew_grid(N * C / 4), 256, 0, (cudaStream_t)stream_>>>(gi_rzn, gh, h, identity, N, C, act, act_param, h_new, x_out);
    else
        gru_gates_fwd_kernel<<<ew_grid(N * C), 256, 0, (cudaStream_t)stream_>>>(gi_rzn, gh, h, identity, N, C, act, act_param, h_new, x_out);
    GLAM_CHECK_LAUNCH();
    return 0;
}

// gh_n: the hidden-side pre-activation of the n gate (W_hn h + b_hn), row n at gh_n + n * ld_ghn
extern "C" int glam_gru_gates_bwd_ex(const float* rzn, const float* gh_n, int64_t ld_ghn, const float* h, const float* x_out,
                                     const float* g_x_out, const float* g_h_carry, int64_t N, int C, int act, float act_param,
                                     float* g_gi, float* g_gh, float* g_h_prev, float* g_identity, void* stream_) {
    GLAM_REQUIRE(N >= 0 && C > 0 && act >= 0 && act <= 3 && ld_ghn >= C, "glam_gru_gates_bwd: bad arguments");
    if (N == 0) return 0;
    GLAM_REQUIRE(rzn && gh_n && h && x_out && g_gi && g_gh && g_h_prev, "glam_gru_gates_bwd: null pointer");
    if ((C & 3) == 0 && (ld_ghn & 3) == 0 && al16(rzn) && al16(gh_n) && al16(h) && al16(x_out) && al16(g_x_out) && al16(g_h_carry) &&
        al16(g_gi) && al16(g_gh) && al16(g_h_prev) && al16(g_identity))
        gru_gates_bwd_vec_kernel<<<ew_grid(N * C / 4), 256, 0, (cudaStream_t)stream_>>>(rzn, gh_n, ld_ghn, h, x_out, g_x_out, g_h_carry, N, C,
                                                                                    act, act_param, g_gi, g_gh, g_h_prev, g_identity);
    else
        gru_gates_bwd_kernel<<<ew_grid(N * C), 256, 0, (cudaStream_t)stream_>>>(rzn, gh_n, ld_ghn, h, x_out, g_x_out, g_h_carry, N, C, act,
                                                                                act_param, g_gi, g_gh, g_h_prev, g_identity);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_gru_gates_bwd(const float* rzn, const float* gh, const float* h, const float* x_out, const float* g_x_out,
                                  const float* g_h_carry, int64_t N, int C, int act, float act_param, float* g_gi,
                                  float* g_gh, float* g_h_prev, float* g_identity, void* stream_) {
    GLAM_REQUIRE(gh != nullptr || N == 0, "glam_gru_gates_bwd: null pointer");
    return glam_gru_gates_bwd_ex(rzn, gh ? gh + 2 * (int64_t)C : nullptr, 3 * (int64_t)C, h, x_out, g_x_out, g_h_carry, N, C, act, act_param,
                                 g_gi, g_gh, g_h_prev, g_identity, stream_);
}

extern "C" int glam_lstm_gates_fwd(float* gates, const float* c_prev, int64_t R, int C, float* c_new, float* h_new, void* stream_) {
    GLAM_REQUIRE(R >= 0 && C > 0, "glam_lstm_gates_fwd: bad arguments");
    if (R == 0) return 0;
    GLAM_REQUIRE(gates && c_prev && c_new && h_new, "glam_lstm_gates_fwd: null pointer");
    lstm_gates_fwd_kernel<<<ew_grid(R * C), 256, 0, (cudaStream_t)stream_>>>(gates, c_prev, R, C, c_new, h_new);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_lstm_gates_bwd(const float* gates_act, const float* c_prev, const float* c_new, const float* g_h,
                                   const float* g_c_in, int64_t R, int C, float* g_gates, float* g_c_prev, void* stream_) {
    GLAM_REQUIRE(R >= 0 && C > 0, "glam_lstm_gates_bwd: bad arguments");
    if (R == 0) return 0;
    GLAM_REQUIRE(gates_act && c_prev && c_new && g_gates && g_c_prev, "glam_lstm_gates_bwd: null pointer");
    lstm_gates_bwd_kernel<<<ew_grid(R * C), 256, 0, (cudaStream_t)stream_>>>(gates_act, c_prev, c_new, g_h, g_c_in, R, C, g_gates, g_c_prev);
    GLAM_CHECK_LAUNCH();
    return 0;
}
