/* xnode_wan_b200.h -- C ABI of the B200-native XNODE-WAN hot path (libxnode_wan_b200.so).
 *
 * Drop-in boundary for ONE path of paulvoliva/XNODE-WAN-PDE-solver: the per-iteration Monte-Carlo
 * weak-form loss and its parameter gradients (reference: src/training.py:128-138 u-phase,
 * :153-162 v-phase).  Plain pointers and sizes only; every pointer is a CUDA DEVICE pointer unless
 * the name ends in _host; `stream` is a cudaStream_t passed as void*.  No call synchronises the
 * host.  Every function returns 0 on success, non-zero on error (xw_last_error() has the text);
 * there is no CPU fallback: a call on a machine without a usable CUDA device fails.
 *
 * Tensor conventions (reference layout [N, L, C], time in channel 0, C = d + 1):
 *   a "path tensor" is described by a base pointer and element strides, so both the reference's
 *   repeated layout (x base = X + 1, x_sn = L*C, x_sl = C) and a collapsed layout
 *   (x[N, d], x_sn = d, x_sl = 0, shared times[L]) are accepted without a copy.
 *
 * Flat parameter layout (fp32, PyTorch-native row-major [out][in], the reference's
 * named_parameters() order):
 *   theta_u = initial_layers.0.{weight[H,1],bias[H]} .2.{weight[H,H],bias[H]} .4.{weight[H,H],bias[H]}
 *             ODE_rhs.net.0.{weight[hh,d+1+H],bias[hh]}  (input order (x, t, y): src/model.py:154)
 *             ODE_rhs.net.2.{weight[hh,hh],bias[hh]}      (shared by the nu-1 hidden layers, src/model.py:130)
 *             ODE_rhs.net.<2nu>.{weight[H,hh],bias[H]}  final_linear.{weight[1,H],bias[1]}
 *   theta_v = input.{weight[Hv,d+1],bias[Hv]} hidden.{weight[Hv,Hv],bias[Hv]} (shared nv times,
 *             src/model.py:39) output.{weight[1,Hv],bias[1]}
 */
#ifndef XNODE_WAN_B200_H
#define XNODE_WAN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XW_ABI_VERSION 4

enum { XW_SOLVER_EULER = 0, XW_SOLVER_MIDPOINT = 1, XW_SOLVER_RK4 = 2 };
enum { XW_DOMAIN_CUBE = 0, XW_DOMAIN_CONE = 1, XW_DOMAIN_HOURGLASS = 2 };

/* network sizes: cube_pde.yaml keys u_hidden_dim, u_hidden_hidden_dim, u_layers, v_hidden_dim,
 * v_layers, solver (src/training.py:94-97, src/model.py:62-63, :30-36) */
typedef struct xw_dims {
    int d, H, hh, nu, Hv, nv, solver;
} xw_dims;

/* domain weight w (src/dataset.py:278-282 cube: p0=bot, p1=top; :216-218 cone: p0=r;
 * :119-125 hourglass: p0=r, p1=T0, p2=T) */
typedef struct xw_domain {
    int kind;
    float p0, p1, p2;
} xw_domain;

/* PDE coefficients as structure instead of the reference's dense tensors a[d,d,N,L], b[d,N,L], c[N,L,1]
 * (src/training.py:25-41).  The common case is constant: c(X,u) = c0 + c1*u; a: NULL = identity, else [d*d] row-major;
 * b: NULL = 0, else [d]; a_sn = b_sn = 0, A_val = A_der = NULL.
 * General callables (the reference accepts any func_a(X,i,j), func_b(X,i), func_c(X,u)):
 *   - a, b that depend on X: only their values on time-row 0 of every path enter the loss (src/loss.py:66-69 multiplies
 *     them with du, which lives on row 0 because the XNODE reads x from row 0, src/model.py:99): pass a[n][d*d] /
 *     b[n][d] per path with a_sn = d*d / b_sn = d (elements between consecutive paths; 0 = one constant matrix / vector);
 *   - c that depends on X or is not affine in u: A_val[n*L] = c(X,u) u and A_der[n*L] = d(c(X,u) u)/du per point,
 *     evaluated by the host with the user's callable at u = xw_xnode_eval(...) (bit-identical to the u the interior
 *     forward computes); when given they replace c0, c1. */
typedef struct xw_coef {
    float c0, c1;
    const float* a;
    const float* b;
    long long a_sn, b_sn;
    const float* A_val;
    const float* A_der;
} xw_coef;

/* strided view of per-point (t, x): t at t[n*t_sn + l*t_sl], x_j at x[n*x_sn + l*x_sl + j] */
typedef struct xw_points {
    const float* t;
    long long t_sn, t_sl;
    const float* x;
    long long x_sn, x_sl;
} xw_points;

/* indices into the `sums` accumulator (double[XW_NSUMS]); all are UNSCALED Monte-Carlo sums:
 *   I    = V/N * sums[S1] - V/(N L) * (sums[S2] - sums[S3])          (src/loss.py:64-73)
 *   S    = V * sums[VV] / (N L)                                       (src/loss.py:90)
 *   init = sums[INIT] / N,  bdry = sums[BDRY] / (N_b L_b)             (src/loss.py:78-85) */
enum { XW_SUM_S1 = 0, XW_SUM_S2 = 1, XW_SUM_S3 = 2, XW_SUM_VV = 3, XW_SUM_INIT = 4, XW_SUM_BDRY = 5, XW_NSUMS = 8 };

int xw_abi_version(void);
const char* xw_last_error(void);
int xw_theta_u_size(const xw_dims* dims);
int xw_theta_v_size(const xw_dims* dims);
/* floats in the optional state-history / test-function cache buffers of xw_interior_forward */
size_t xw_yhist_floats(const xw_dims* dims, int n, int L);
/* which generation of the XNODE kernels this process launched last: bits 0-3 = the last XNODE launch of any kind,
 * bits 4-7 = the last BACKWARD launch (1 = one thread per path, 2 = reduced state / 2 warp roles, 3 = 3 warp roles;
 * 0 = none yet): lets tests and bench.py name the kernel that actually ran instead of assuming it */
int xw_last_xnode_impl(void);
/* same for the test-function net: bits 0-3 = last forward launch, bits 4-7 = last backward launch;
 * 1 = one thread per point (FP32), 2 = FP32 tile engine, 3 = tcgen05 (3xTF32), 4 = tcgen05 on the virtual net of
 * input width Hv (backward for d > 54) */
int xw_last_vnet_impl(void);
size_t xw_vcache_floats(const xw_dims* dims, int n, int L);
/* upper bound of the scratch any call below needs for n paths of length L */
size_t xw_workspace_bytes(const xw_dims* dims, int n, int L);

/* u = u_net(X) forward only -> u_out[n*L]   (replaces NeuralODE.forward, src/model.py:87-112).
 * x: spatial coords of time-row 0 (x_j at x[i*x_sn + j]); times[L]: path 0's time grid
 * (src/model.py:92); s0[n]: initial scalar h(x) or g(x) (src/model.py:95-96). */
int xw_xnode_eval(const xw_dims* dims, const float* theta_u, const float* x, long long x_sn,
                  const float* times, int L, const float* s0, int n, float* u_out, void* stream);

/* v = v_net(XV) forward only -> v_out[n*L]   (replaces discriminator.forward, src/model.py:45-47) */
int xw_vnet_eval(const xw_dims* dims, const float* theta_v, const xw_points* pts, int n, int L,
                 float* v_out, void* stream);

/* Interior forward pass (replaces v_net(XV), u_net(X) and loss.I / loss.init / the S term:
 * src/training.py:129-130, src/loss.py:46-80,87-90).  Adds this shard's contributions to
 * sums[S1,S2,S3,VV,INIT]; writes the per-point cotangent seeds the backward passes need:
 *   cot_u[i,l] = A'(u) phi + L v[i,L-1] [l=L-1]                 A(u) = (c0 + c1 u) u
 *   cot_v[i,l] = w (A(u) + f) + L u[i,L-1] [l=L-1] - L h[i] [l=0]
 * and optionally u_out[n*L] (NULL to skip).
 * h[n] = func_h(X[:,0,:]) (src/training.py:25), f[n*L] = func_f(X): the user's callables evaluated
 * by the host.  s0[n]: initial scalar of each path (src/model.py:95-96: h(x) when the batch starts
 * at T0, g at the entry point otherwise; NULL = use h); grad_s0[n*d] = d s0 / d x.
 * vcache (xw_vcache_floats(dims, n, L) floats) / vcache_mode: the reference runs n1 u-steps and n2
 * v-steps on ONE sample (src/training.py:125,151); while the sample and theta_v are unchanged the
 * test-function values (v, dv/dt, w, dw/dt per point, grad_x phi on time-row 0) do not change either:
 *   0 = no cache, 1 = evaluate the v net and fill the cache, 2 = skip the v net and read the cache.
 * y_hist (optional, xw_yhist_floats(dims, n, L) floats): the XNODE state history y_l of every path, kept so
 * that xw_interior_backward_u does not integrate the ODE forward again. */
int xw_interior_forward(const xw_dims* dims, const xw_domain* dom, const xw_coef* coef,
                        const float* theta_u, const float* theta_v,
                        const float* x, long long x_sn, const float* times, int L,
                        const xw_points* xv, const float* h, const float* grad_s0, const float* f,
                        int n, double* sums, float* cot_u, float* cot_v, float* u_out,
                        void* workspace, size_t workspace_bytes, void* stream, const float* s0,
                        float* vcache, int vcache_mode, float* y_hist, size_t vcache_floats, size_t y_hist_floats);
/* vcache_floats / y_hist_floats: capacities of the two optional buffers in floats; the call fails (no write) when a
 * buffer that is passed is smaller than xw_vcache_floats / xw_yhist_floats for this (n, L). */

/* Boundary term (replaces loss.bdry = mean((u_net(BX) - g)^2), src/loss.py:83-85, and its
 * backward): adds sum (u_b - g)^2 to sums[BDRY]; if grad_u != NULL also accumulates
 * d/dtheta_u of  gscale * sum (u_b - g)^2  (gscale = alpha / (N_b L_b), global counts). */
int xw_boundary_u(const xw_dims* dims, const float* theta_u, const float* xb, long long xb_sn,
                  const float* times_b, int Lb, const float* s0b, const float* g, int nb,
                  double gscale, double* sums, float* grad_u, int accumulate,
                  void* workspace, size_t workspace_bytes, void* stream);

/* theta_u gradient of the interior part of loss_u (replaces loss_u.backward() for u_net,
 * src/training.py:137, including the reference's side effect of src/loss.py:55):
 *   G_u[i,l] = k[0]*cot_u[i,l] + k[1]*(u[i,0]-h[i])[l=0] + k[2]
 * k = coefs_dev[0..2] (DEVICE doubles, so no host sync): k0 = (2/I) V/(N L), k1 = 2 alpha/N, k2 = 1 */
int xw_interior_backward_u(const xw_dims* dims, const float* theta_u, const float* x, long long x_sn,
                           const float* times, int L, const float* h, const float* cot_u, int n,
                           const double* coefs_dev, float* grad_u, int accumulate,
                           void* workspace, size_t workspace_bytes, void* stream, const float* s0,
                           const float* y_hist);

/* theta_v gradient of loss_v (replaces loss_v.backward() for v_net, src/training.py:160,
 * including the side effect of src/loss.py:60):
 *   G_v[i,l] = k[0]*cot_v[i,l] + k[1]*v[i,l] + k[2]*w[i,l]
 * k0 = -(2/I) V/(N L), k1 = 2/sums[VV], k2 = 1 */
int xw_interior_backward_v(const xw_dims* dims, const xw_domain* dom, const float* theta_v,
                           const xw_points* xv, const float* cot_v, int n, int L,
                           const double* coefs_dev, float* grad_v, int accumulate,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Adam step of one net on its flat parameter vector (reference: torch.optim.Adam(net.parameters(), lr) with default
 * betas / eps, /root/reference/src/training.py:103-104, stepped at :138 / :162).  params / exp_avg / exp_avg_sq: fp64 [n];
 * grad: the flat fp32 gradient the backward entries produce; step: device counter (incremented by the call); params_f32
 * (optional): receives the updated parameters in fp32 = the theta_u / theta_v argument of the next forward call. */
int xw_adam_step(double* params, const float* grad, double* exp_avg, double* exp_avg_sq, long long* step,
                 float* params_f32, int n, double lr, double beta1, double beta2, double eps, void* stream);

/* The scalars of one sub-step from the (all-reduced) sums, in fp64 on the device, one launch (ABI v4).  Replaces the
 * arithmetic of /root/reference/src/loss.py:64-76 (I, S), :78-90 (init, bdry, int) and :92-96 (loss u / v) on the eight
 * sums, and forms the coefficients `coefs_dev` of the backward entries above:
 *   out[0] = loss: phase 0 (u): log(I^2) - log(S) + alpha (init + bdry);  phase 1 (v): -(log(I^2) - log(S))
 *   out[1..4] = I, S, init, bdry          out[5..7] = k0, k1, k2
 * n_glob, nb_glob: GLOBAL path counts (all ranks); nb_glob = 0: no boundary batch (bdry = 0). */
int xw_loss_scalars(const double* sums, int phase, double V, double n_glob, double L, double nb_glob, double Lb,
                    double alpha, double side, double* out, void* stream);

/* FP32-FMA micro-benchmark used as the roofline denominator: runs `iters`
 * dependent-chain FFMA blocks on every SM and returns the FLOP count in *flops_host. */
int xw_fma_probe(int variant, int iters, double* flops_host, void* stream);

/* tcgen05 self-check (groundwork for the tensor-core variant of the Hv x Hv contractions):
 * D[128 x N] = A[128 x K] * B[N x K]^T with kind::tf32 MMAs, accumulator in TMEM; terms = 1 plain TF32,
 * 3 = 3xTF32 error compensation.  K multiple of 8 (<= 64), N multiple of 16 (<= 64). */
int xw_umma_probe(const float* A, const float* B, float* D, int K, int N, int terms, int* err_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
