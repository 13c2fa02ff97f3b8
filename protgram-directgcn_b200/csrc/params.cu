// Hot path B, parameter side of one DirectGCN layer (reference src/models/protgram_directgcn.py:34-66
// parameters, :101-133 how they enter the layer).  The reference applies 4 Linears, 6 biases and 5
// gate vectors as ~25 separate tensor ops per layer and direction; the fused layer needs them as
//   W_ext = [(W_in+W_sh)^T; (W_out+W_sh)^T; (W_und+W_sh)^T; W_res^T?; b_in+bs_in; b_out+bs_out; b_und+bs_und; b_res?]
//   a = (C_all*C_dir)*C_in,  b = (C_all*C_dir)*C_out,  c = C_all*C_und
// One kernel packs them (same fp32 operations, so bit-identical to the tensor-op composition), one
// kernel scatters dW_ext / da / db / dc back onto the reference's parameters.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) pack_kernel(pg_layer_params p, int64_t num_gate, int F_in, int F_out, int has_res,
                                                   float *__restrict__ w_ext, float *__restrict__ ga, float *__restrict__ gb,
                                                   float *__restrict__ gc) {
    const int k_data = 3 * F_in + (has_res ? F_in : 0);
    const int k_ext = k_data + 3 + (has_res ? 1 : 0);
    const int64_t n_w = (int64_t)k_ext * F_out;
    const int64_t total = n_w + num_gate;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        if (t < n_w) {
            const int k = (int)(t / F_out), o = (int)(t % F_out);
            float v;
            if (k < 3 * F_in) {
                const int m = k / F_in, i = k - m * F_in;
                const float *w = m == 0 ? p.w_in : (m == 1 ? p.w_out : p.w_und);
                v = __fadd_rn(w[(int64_t)o * F_in + i], p.w_sh[(int64_t)o * F_in + i]);
            } else if (k < k_data) {
                v = p.w_res[(int64_t)o * F_in + (k - 3 * F_in)];
            } else {
                const int j = k - k_data;
                if (j == 0) v = __fadd_rn(p.b_in[o], p.bs_in[o]);
                else if (j == 1) v = __fadd_rn(p.b_out[o], p.bs_out[o]);
                else if (j == 2) v = __fadd_rn(p.b_und[o], p.bs_und[o]);
                else v = p.b_res ? p.b_res[o] : 0.f;
            }
            w_ext[t] = v;
        } else {
            const int64_t g = t - n_w;
            const float call = p.c_all[g];
            const float cd = __fmul_rn(call, p.c_dir[g]);
            ga[g] = __fmul_rn(cd, p.c_in[g]);
            gb[g] = __fmul_rn(cd, p.c_out[g]);
            gc[g] = __fmul_rn(call, p.c_und[g]);
        }
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(pg_layer_params p, pg_layer_param_grads d, const float *__restrict__ dw_ext,
                                                     const float *__restrict__ dga, const float *__restrict__ dgb,
                                                     const float *__restrict__ dgc, int64_t num_gate, int F_in, int F_out,
                                                     int has_res) {
    const int k_data = 3 * F_in + (has_res ? F_in : 0);
    const int64_t n_lin = (int64_t)F_out * F_in;
    const int64_t total = n_lin + F_out + num_gate;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        if (t < n_lin) {  // [o][i], i fastest: the parameters' own layout
            const int o = (int)(t / F_in), i = (int)(t % F_in);
            const float a = dw_ext[(int64_t)i * F_out + o];
            const float b = dw_ext[(int64_t)(F_in + i) * F_out + o];
            const float c = dw_ext[(int64_t)(2 * F_in + i) * F_out + o];
            d.w_in[t] = a;
            d.w_out[t] = b;
            d.w_und[t] = c;
            d.w_sh[t] = (a + b) + c;
            if (has_res) d.w_res[t] = dw_ext[(int64_t)(3 * F_in + i) * F_out + o];
        } else if (t < n_lin + F_out) {
            const int o = (int)(t - n_lin);
            const float a = dw_ext[(int64_t)k_data * F_out + o];
            const float b = dw_ext[(int64_t)(k_data + 1) * F_out + o];
            const float c = dw_ext[(int64_t)(k_data + 2) * F_out + o];
            d.b_in[o] = a; d.bs_in[o] = a;
            d.b_out[o] = b; d.bs_out[o] = b;
            d.b_und[o] = c; d.bs_und[o] = c;
            if (has_res && d.b_res) d.b_res[o] = dw_ext[(int64_t)(k_data + 3) * F_out + o];
        } else {
            const int64_t g = t - n_lin - F_out;
            const float call = p.c_all[g], cdir = p.c_dir[g], cin = p.c_in[g], cout = p.c_out[g], cund = p.c_und[g];
            const float da = dga[g], db = dgb[g], dc = dgc[g];
            const float cd = call * cdir;
            const float s = da * cin + db * cout;   // d(loss)/d(C_all*C_dir)
            d.c_in[g] = da * cd;
            d.c_out[g] = db * cd;
            d.c_dir[g] = s * call;
            d.c_und[g] = dc * call;
            d.c_all[g] = s * cdir + dc * cund;
        }
    }
}

inline unsigned grid_for(int64_t n) {
    int64_t want = pg_ceil_div(n, 256);
    const int64_t cap = (int64_t)PG_NUM_SMS * 8;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

bool params_ok(const pg_layer_params *p, int has_res) {
    return p && p->w_in && p->w_out && p->w_und && p->w_sh && p->b_in && p->b_out && p->b_und && p->bs_in && p->bs_out &&
           p->bs_und && p->c_in && p->c_out && p->c_dir && p->c_und && p->c_all && (!has_res || p->w_res);
}
}  // namespace

extern "C" int pg_pack_layer_params(const pg_layer_params *params, int64_t num_gate, int F_in, int F_out, int has_res,
                                    float *d_w_ext, float *d_gate_a, float *d_gate_b, float *d_gate_c, pg_stream_t stream) {
    PG_CHECK_ARG(F_in >= 1 && F_out >= 1 && num_gate >= 1, "pg_pack_layer_params: bad shape");
    PG_CHECK_ARG(params_ok(params, has_res), "pg_pack_layer_params: null parameter pointer");
    PG_CHECK_ARG(d_w_ext && d_gate_a && d_gate_b && d_gate_c, "pg_pack_layer_params: null output");
    const int64_t total = (int64_t)(3 * F_in + (has_res ? F_in + 1 : 0) + 3) * F_out + num_gate;
    pack_kernel<<<grid_for(total), 256, 0, pg_cu(stream)>>>(*params, num_gate, F_in, F_out, has_res ? 1 : 0, d_w_ext, d_gate_a,
                                                             d_gate_b, d_gate_c);
    PG_CUDA_LAUNCH_CHECK("pack_kernel");
    return PG_OK;
}

extern "C" int pg_unpack_layer_param_grads(const pg_layer_params *params, const float *d_dw_ext, const float *d_dgate_a,
                                           const float *d_dgate_b, const float *d_dgate_c, int64_t num_gate, int F_in, int F_out,
                                           int has_res, const pg_layer_param_grads *grads, pg_stream_t stream) {
    PG_CHECK_ARG(F_in >= 1 && F_out >= 1 && num_gate >= 1, "pg_unpack_layer_param_grads: bad shape");
    PG_CHECK_ARG(params_ok(params, 0) && d_dw_ext && d_dgate_a && d_dgate_b && d_dgate_c, "pg_unpack_layer_param_grads: null input");
    const pg_layer_param_grads *g = grads;
    PG_CHECK_ARG(g && g->w_in && g->w_out && g->w_und && g->w_sh && g->b_in && g->b_out && g->b_und && g->bs_in && g->bs_out &&
                     g->bs_und && g->c_in && g->c_out && g->c_dir && g->c_und && g->c_all && (!has_res || g->w_res),
                 "pg_unpack_layer_param_grads: null gradient pointer");
    const int64_t total = (int64_t)F_out * F_in + F_out + num_gate;
    unpack_kernel<<<grid_for(total), 256, 0, pg_cu(stream)>>>(*params, *grads, d_dw_ext, d_dgate_a, d_dgate_b, d_dgate_c, num_gate,
                                                               F_in, F_out, has_res ? 1 : 0);
    PG_CUDA_LAUNCH_CHECK("unpack_kernel");
    return PG_OK;
}
