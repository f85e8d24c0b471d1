// chain_tma_kernel.cu -- the serial-chain kernel (chain_kernel.cu) with its outputs leaving the SM through the TMA unit.
//
// Same arithmetic as exo_chain_kernel (modified-DH chain in registers, KinematicTree::ComputeJ columns, the task maps' Phi
// encodings; reference: exotica_core/src/kinematic_tree.cpp:678-737,829-867,919-945, eff_frame.cpp:53-100,
// eff_position.cpp:36-75), different store path: one CTA = one tile of TILE evaluations; every lane writes its Phi / J
// entries into a shared-memory tile [row][TILE] with immediate-offset st.shared (the per-lane st.global path spends 21 % of its
// warp instructions on 64-bit address arithmetic), then one lane issues one cp.async.bulk.tensor.2d per block of rows (L2
// evict-first).  The hardware clips the ragged last tile.  JointLimit rows / unused-map fills keep the per-lane streaming stores
// of chain_prefix.
//
// Measured on B200 (cfg2a, 2^20 evaluations, tools/bench_configs.py; bare patterns in tools/tma_probe.cu, profiles/r02_tma_probe.txt):
//   per-lane st.global.cs, 64-lane CTAs (exo_chain_kernel)          85.4 us
//   this kernel, 64-lane tiles (144 registers, 7 CTAs / SM)           85.6 us
//   this kernel, 128-lane tiles (128 registers, 4 CTAs / SM)          83.6 us   <- default where it applies
//   long-lived CTAs walking several tiles as a cp.async / TMA pipeline (joint values of the next tile in flight, staging tile
//   reused one tile later; 1 / 2 / 4 / 8 / 16 tiles per CTA or one CTA per resident slot): 102 / 98 / 96 / 96 / 96 / 94 us --
//   slower at every setting (barrier stalls 1.4 per issue, the next tile's values still not there after a whole tile of
//   arithmetic: reads queue behind the write-back of a cache full of dirty lines), so one tile per CTA + L2 prefetch stays.
//   EffPosition-only and JointLimit-carrying problems (cfg2b, cfg3) are 10-15 % slower through the staging tile than with
//   per-lane stores (less data per barrier; mixed TMA / per-lane stores), so the launcher keeps them on exo_chain_kernel.
#include <stdlib.h>
#include <string.h>

#include "chain_device.cuh"
#include "tma.cuh"

namespace exo
{
namespace
{
using namespace dev;

// Jacobian column of joint J for one requested frame into the staging tile: sj = &stage[first J row of the frame][lane];
// row r of column J sits NJ * r + J rows further (the frame's block is [n_rows][N] with N == NJ, column index == joint index)
template <int J, int NJ, int ALLREV, int TILE, bool WITHB, bool SIX>
__device__ __forceinline__ void stage_column(const ChainParams& P, const ChainFrameP& fp, double* sj, double wx, double wy, double wz, double px, double py,
                                             double pz, double ax, double ay, double az)
{
    const ChainJointP& jp = P.joint[J];
    double v0, v1, v2, w0, w1, w2;
    if (!ALLREV && jp.prismatic)
    {
        v0 = wx; v1 = wy; v2 = wz;
        w0 = w1 = w2 = 0.0;
    }
    else
    {
        const double dx = ax - px, dy = ay - py, dz = az - pz;
        v0 = wy * dz - wz * dy;
        v1 = wz * dx - wx * dz;
        v2 = wx * dy - wy * dx;
        w0 = wx; w1 = wy; w2 = wz;
    }
    if (WITHB)
    {
        const double t0 = fp.binv[0] * v0 + fp.binv[1] * v1 + fp.binv[2] * v2;
        const double t1 = fp.binv[3] * v0 + fp.binv[4] * v1 + fp.binv[5] * v2;
        const double t2 = fp.binv[6] * v0 + fp.binv[7] * v1 + fp.binv[8] * v2;
        v0 = t0; v1 = t1; v2 = t2;
        if (SIX)
        {
            const double u0 = fp.binv[0] * w0 + fp.binv[1] * w1 + fp.binv[2] * w2;
            const double u1 = fp.binv[3] * w0 + fp.binv[4] * w1 + fp.binv[5] * w2;
            const double u2 = fp.binv[6] * w0 + fp.binv[7] * w1 + fp.binv[8] * w2;
            w0 = u0; w1 = u1; w2 = u2;
        }
    }
    sj[(0 * NJ + J) * TILE] = v0;
    sj[(1 * NJ + J) * TILE] = v1;
    sj[(2 * NJ + J) * TILE] = v2;
    if (SIX)
    {
        sj[(3 * NJ + J) * TILE] = w0;
        sj[(4 * NJ + J) * TILE] = w1;
        sj[(5 * NJ + J) * TILE] = w2;
    }
}

template <int NJ, int ALLREV, int TILE, int J>
struct StageColumns
{
    template <bool WITHB, bool SIX>
    static __device__ __forceinline__ void run(const ChainParams& P, const ChainFrameP& fp, double* sj, const double (&W)[NJ][3], const double (&PJ)[NJ][3],
                                               double ax, double ay, double az)
    {
        stage_column<J, NJ, ALLREV, TILE, WITHB, SIX>(P, fp, sj, W[J][0], W[J][1], W[J][2], PJ[J][0], PJ[J][1], PJ[J][2], ax, ay, az);
        StageColumns<NJ, ALLREV, TILE, J + 1>::template run<WITHB, SIX>(P, fp, sj, W, PJ, ax, ay, az);
    }
};
template <int NJ, int ALLREV, int TILE>
struct StageColumns<NJ, ALLREV, TILE, NJ>
{
    template <bool WITHB, bool SIX>
    static __device__ __forceinline__ void run(const ChainParams&, const ChainFrameP&, double*, const double (&)[NJ][3], const double (&)[NJ][3], double,
                                               double, double) {}
};

#ifndef EXO_CHAIN_TMA_MAXNREG64
#define EXO_CHAIN_TMA_MAXNREG64 144
#endif
#ifndef EXO_CHAIN_TMA_MAXNREG128
#define EXO_CHAIN_TMA_MAXNREG128 128
#endif

template <int NJ, int ALLREV, int TILE>
__global__ void __launch_bounds__(TILE) __maxnreg__(TILE == 64 ? EXO_CHAIN_TMA_MAXNREG64 : EXO_CHAIN_TMA_MAXNREG128)
    exo_chain_tma_kernel(const __grid_constant__ ChainParams P, const __grid_constant__ ChainTmaParams M)
{
    extern __shared__ __align__(128) double stage[];  // [M.stage_rows][TILE]
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * TILE;
    const bool valid = e0 + tid < P.n_evals;
    // lanes beyond the batch recompute the last evaluation into their own tile column; the TMA unit clips that column
    const int64_t e = valid ? e0 + tid : P.n_evals - 1;
    const char* xe = reinterpret_cast<const char*>(P.x) + e * P.x_se_b;
    double q[NJ];
#pragma unroll
    // joint values are read once: streaming loads (evict-first in L2, which is full of the outputs' dirty lines): 82.3 -> 81.8 us on cfg2a.
    // Not in exo_chain_kernel: its JointLimit prefix reads the state a second time (cfg3 67.7 -> 72.1 us with streaming loads there).
    for (int j = 0; j < NJ; ++j) q[j] = __ldcs(reinterpret_cast<const double*>(xe + P.joint[j].x_byte_off));
    // as in exo_chain_kernel: pull the joint values of a CTA most of a resident generation behind this one into L2
    if (e + P.prefetch_ahead < P.n_evals)
    {
#pragma unroll
        for (int j = 0; j < NJ; ++j) prefetch_l2(xe + P.prefetch_ahead * P.x_se_b + P.joint[j].x_byte_off);
    }
    if (P.has_prefix && valid)
    {
        Sink o;
        o.phi = P.phi + e;
        o.jac = P.jac ? P.jac + e : nullptr;
        o.phi_sk = P.phi_sk;
        o.j_sr = P.j_sr;
        o.j_sc = P.j_sc;
        chain_prefix(P, o, reinterpret_cast<const double*>(xe), P.x_sk);
    }
    double sn[NJ], cs[NJ];
    bool big = false;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
    {
        const double th = fma(P.joint[j].qmult, q[j], P.joint[j].th0);
        big = big || !(fabs(th) <= 1.0e5);
        sn[j] = 0.0;
        cs[j] = 1.0;
        if (ALLREV || !P.joint[j].prismatic) fast_sincos(th, sn[j], cs[j]);
    }
    if (__syncthreads_or(big))
    {
        // some angle of this tile is beyond the fast reduction: the whole tile takes the cold per-lane path (CUDA's sincos)
        if (valid) chain_slow_eval(P, xe, reinterpret_cast<char*>(P.phi + e), P.jac ? reinterpret_cast<char*>(P.jac + e) : nullptr);
        return;
    }
    Fr cur = load_frame(P.base);
    double W[NJ][3], PJ[NJ][3];
    ChainUnroll<NJ, ALLREV, 0>::record(P, cur, q, sn, cs, W, PJ);
    double* const lane = stage + tid;
    if (P.jac)
    {
        for (int f = 0; f < P.n_frames; ++f)
        {
            const ChainFrameP& fp = P.frame[f];
            const Fr A = chain_frame_a(fp, cur);
            double* sj = lane + M.stage_j_row[f] * TILE;
            const bool six = fp.row_mode == ROWS_VW;
            if (fp.b_const)
            {
                if (six) StageColumns<NJ, ALLREV, TILE, 0>::template run<true, true>(P, fp, sj, W, PJ, A.x, A.y, A.z);
                else StageColumns<NJ, ALLREV, TILE, 0>::template run<true, false>(P, fp, sj, W, PJ, A.x, A.y, A.z);
            }
            else
            {
                if (six) StageColumns<NJ, ALLREV, TILE, 0>::template run<false, true>(P, fp, sj, W, PJ, A.x, A.y, A.z);
                else StageColumns<NJ, ALLREV, TILE, 0>::template run<false, false>(P, fp, sj, W, PJ, A.x, A.y, A.z);
            }
        }
    }
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        Fr A = chain_frame_a(fp, cur);
        if (fp.b_const) A = fr_mul(load_frame(fp.binv), A);
        write_phi(fp.map_kind, fp.rot_type, A, lane + M.stage_phi_row[f] * TILE, TILE);
    }
    tma::fence_staged();
    __syncthreads();
    if (tid == 0)
    {
        const uint64_t pol = tma::policy_evict_first();
        for (int f = 0; f < P.n_frames; ++f)
        {
            if (P.jac) tma::store_2d(&M.jac[f], stage + M.stage_j_row[f] * TILE, (int)e0, 0, pol);
            tma::store_2d(&M.phi[f], stage + M.stage_phi_row[f] * TILE, (int)e0, 0, pol);
        }
        tma::commit();
        tma::wait_read_all();
    }
}

int g_sms = 0;
int prefetch_quarter_generations()
{
    static int gens = 0;
    if (gens == 0)
    {
        const char* env = getenv("EXO_CHAIN_PF_QGENS");
        gens = env ? atoi(env) : 3;
        if (gens < 1) gens = 1;
    }
    return gens;
}

template <int NJ, int ALLREV, int TILE>
cudaError_t tma_launch_t(const ChainParams& P, const ChainTmaParams& M, cudaStream_t stream)
{
    static int ctas_per_sm = 0;
    static size_t smem_set = 0;
    const size_t smem = (size_t)M.stage_rows * TILE * sizeof(double);
    auto kernel = exo_chain_tma_kernel<NJ, ALLREV, TILE>;
    if (smem > smem_set)
    {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        smem_set = smem;
        ctas_per_sm = 0;
    }
    if (ctas_per_sm == 0)
    {
        int ctas = 0;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, TILE, smem);
        if (err != cudaSuccess) return err;
        if (ctas < 1) return cudaErrorInvalidConfiguration;
        ctas_per_sm = ctas;
    }
    const int64_t blocks = (P.n_evals + TILE - 1) / TILE;
    if (blocks > 0x7fffffff) return cudaErrorInvalidConfiguration;
    ChainParams Q = P;
    Q.prefetch_ahead = (int64_t)g_sms * ctas_per_sm * TILE * prefetch_quarter_generations() / 4;
    kernel<<<(int)blocks, TILE, smem, stream>>>(Q, M);
    return cudaGetLastError();
}

template <int ALLREV, int TILE>
cudaError_t tma_launch_nj(const ChainParams& P, const ChainTmaParams& M, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return tma_launch_t<1, ALLREV, TILE>(P, M, stream);
        case 2: return tma_launch_t<2, ALLREV, TILE>(P, M, stream);
        case 3: return tma_launch_t<3, ALLREV, TILE>(P, M, stream);
        case 4: return tma_launch_t<4, ALLREV, TILE>(P, M, stream);
        case 5: return tma_launch_t<5, ALLREV, TILE>(P, M, stream);
        case 6: return tma_launch_t<6, ALLREV, TILE>(P, M, stream);
        case 7: return tma_launch_t<7, ALLREV, TILE>(P, M, stream);
        case 8: return tma_launch_t<8, ALLREV, TILE>(P, M, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

template <int TILE>
cudaError_t tma_launch_tile(const ChainParams& P, const ChainTmaParams& M, cudaStream_t stream)
{
    if (P.all_revolute && P.ortho) return tma_launch_nj<3, TILE>(P, M, stream);
    return P.all_revolute ? tma_launch_nj<1, TILE>(P, M, stream) : tma_launch_nj<0, TILE>(P, M, stream);
}

// Phi entries of one requested frame (write_phi, device_math.cuh)
int frame_phi_rows(const ChainFrameP& fp)
{
    switch (fp.map_kind)
    {
        case EXO_MAP_FRAMES: return 12;
        case EXO_MAP_EFF_POSITION: return 3;
        case EXO_MAP_EFF_FRAME:
        {
            const int rl = rotation_length(fp.rot_type);
            return rl < 0 ? -1 : 3 + rl;
        }
        default: return -1;
    }
}
}  // namespace

bool chain_tma_plan(const ChainParams& P, int tile, ChainTmaParams& M)
{
    static int disabled = -1;
    if (disabled < 0)
    {
        const char* env = getenv("EXO_CHAIN_TMA");
        disabled = (env && atoi(env) == 0) ? 1 : 0;
    }
    if (disabled && tile == 0) return false;
    // layout: evaluations contiguous, 16-byte row pitches, J rows flattened [LJ * N][E]
    if (P.n_evals < 1 || P.n_evals > 0x7fffffff) return false;
    if (P.phi_se_b != 8 || (P.jac && P.j_se_b != 8)) return false;
    if (P.phi_sk < P.n_evals || (P.phi_sk & 1)) return false;
    if (P.jac && (P.j_sc < P.n_evals || (P.j_sc & 1) || P.j_sr != (int64_t)P.n_ctrl * P.j_sc)) return false;
    // every controlled joint is a chain joint, in column order
    if (P.n_ctrl != P.n_joints || P.zero_cols != 0 || P.n_frames < 1) return false;
    for (int j = 0; j < P.n_joints; ++j)
        if (P.joint[j].col != j) return false;
    int rows = 0;
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        if (fp.row_mode != ROWS_VW && fp.row_mode != ROWS_V) return false;
        if (fp.n_rows != (fp.row_mode == ROWS_VW ? 6 : 3)) return false;
        const int pr = frame_phi_rows(fp);
        if (pr < 0) return false;
        M.stage_phi_row[f] = rows;
        rows += pr;
        M.stage_j_row[f] = rows;
        if (P.jac) rows += fp.n_rows * P.n_joints;
        if (fp.n_rows * P.n_joints > 256) return false;
    }
    M.stage_rows = rows;
    if (tile == 0)
    {
        // automatic choice: only where the staging tile measured faster than per-lane stores (six-row EffFrame blocks, no per-lane
        // prefix stores beside them; see the table at the top of this file)
        if (P.has_prefix || !P.jac) return false;
        for (int f = 0; f < P.n_frames; ++f)
            if (P.frame[f].row_mode != ROWS_VW) return false;
        const char* env = getenv("EXO_CHAIN_TMA");
        tile = env && atoi(env) == 64 ? 64 : 128;
    }
    if (tile != 64 && tile != 128) return false;
    if ((size_t)rows * tile * sizeof(double) > 200 * 1024) tile = 64;
    if ((size_t)rows * tile * sizeof(double) > 200 * 1024) return false;
    M.tile = tile;
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        if (!tma_encode_2d(M.phi[f], P.phi + (int64_t)fp.phi_off * P.phi_sk, (uint64_t)P.n_evals, (uint64_t)frame_phi_rows(fp), (uint64_t)P.phi_sk * 8,
                           (uint32_t)tile, (uint32_t)frame_phi_rows(fp)))
            return false;
        if (P.jac && !tma_encode_2d(M.jac[f], P.jac + (int64_t)fp.jrow_off * P.j_sr, (uint64_t)P.n_evals, (uint64_t)fp.n_rows * P.n_joints,
                                    (uint64_t)P.j_sc * 8, (uint32_t)tile, (uint32_t)(fp.n_rows * P.n_joints)))
            return false;
    }
    return true;
}

cudaError_t chain_tma_launch(const ChainParams& P, const ChainTmaParams& M, cudaStream_t stream)
{
    if (g_sms == 0)
    {
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        err = cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return err;
    }
    return M.tile == 64 ? tma_launch_tile<64>(P, M, stream) : tma_launch_tile<128>(P, M, stream);
}

}  // namespace exo
