// kernels.cu -- CUDA interpreter for the compiled tree program (sm_100a).
//
// Replaces, for a whole batch at once, the reference's per-configuration
//   KinematicTree::Update -> UpdateTree -> UpdateFK -> UpdateJ   (exotica_core/src/kinematic_tree.cpp:678-737,829-867,919-988)
//   EffFrame/EffPosition/JointLimit/SphereCollision::Update      (exotations/exotica_core_task_maps/src/*.cpp)
//   Phi.SetZero / jacobian.setZero of the problem Update           (abstract_time_indexed_problem.cpp:243-244)
//
// Mapping: one lane = one evaluation (configuration or (trajectory,timestep)); a warp walks the op list in
// lock-step, so every branch on op fields is warp-uniform and every program constant is a shared-memory
// broadcast.  The running frame lives in registers (FP64); frames/joint data that must outlive the next op
// sit in per-thread shared-memory slots laid out [slot word][thread] (conflict-free).  The kernel is
// persistent (grid = SMs x resident CTAs, grid-stride over the batch) and writes every entry of Phi and J
// exactly once (zeros included), so no memset pass is needed.  HBM-bound FP64 streaming work: no tensor cores.
#include "kernels.cuh"

#include "device_math.cuh"

namespace exo
{
namespace
{

using namespace dev;

// ---- SphereCollision::Update (sphere_collision.cpp:123-162), one evaluation at a time by a whole warp ----------------------
// phi[g] = sum over sphere pairs (i in group A, j in group B) of 1 / (1 + exp(5 eps (|p_i - p_j| - r_i - r_j)));
// J.row(g) = - sum over pairs of (J_i - J_j)^T u_ij,  u_ij = (p_i - p_j) / |p_i - p_j|   (no sigmoid factor: App. C #4).
// Regrouped per sphere:  J.row(g)[c] = - sum_i (w_c x (p_i - p_c)) . g_i + sum_j (w_c x (p_j - p_c)) . h_j  with
// g_i = sum_j u_ij, h_j = sum_i u_ij, over the spheres that hang below joint c.  Each lane owns one sphere of the
// "owner" group and a strided share of the other group; sums over lanes are warp-shuffle reductions.
__constant__ double kExp[14] = {
    1.4426950408889634074,      // 0: log2(e)
    6.93147180369123816490e-01, // 1: ln2 high part (fdlibm)
    1.90821492927058770002e-10, // 2: ln2 low part
    // 3..12: degree-9 interpolant of exp at the Chebyshev nodes of |r| <= ln2 / 2, highest power first (max relative error
    // 1.9e-14, i.e. < 5e-15 absolute on the sigmoid; two FMAs fewer than the degree-11 Taylor sum it replaces)
    2.76326406745051741403e-06, // c9
    2.48844597769495494779e-05, // c8
    1.98411906475530279946e-04, // c7
    1.38888017496096465887e-03, // c6
    8.33333336731158572097e-03, // c5
    4.16666670405521702425e-02, // c4
    1.66666666666156482179e-01, // c3
    4.99999999994385879720e-01, // c2
    1.00000000000000133227e+00, // c1
    1.00000000000001354472e+00, // c0
    0.0};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exp(t) for the sigmoid: |relative error| < 2e-14 (degree-9 Chebyshev interpolant on |r| <= ln2/2).  Instead of clamping t (an FP64 min and an
// FP64 max: a dozen instructions each way on this machine) the power of two is clamped as an integer: beyond 2^+-1000 the sigmoid
// is 0 or 1 to 300 digits either way.  Valid for |t| < 2^31 ln 2 (t = 5 eps x a surface distance in metres).
// Measured and dropped: a table-driven form (2^(j/32) from a 32-entry table in L1, degree-4 polynomial on |r| <= ln2 / 64: ten FP64
// instructions instead of thirteen, same accuracy) -- cfg5 1.93 ms against 1.85 ms: the lookup and the integer work sit in the
// dependent chain of every pair and cost more issue slots than the three FMAs they save.
__device__ __forceinline__ double sph_exp(double t)
{
    const double magic = 6755399441055744.0;
    const double kf = fma(t, kExp[0], magic);
    const int k = max(min(__double2loint(kf), 1000), -1000);
    const double kd = kf - magic;
    double r = fma(-kd, kExp[1], t);
    r = fma(-kd, kExp[2], r);
    double p = fma(r, kExp[3], kExp[4]);
    p = fma(r, p, kExp[5]);
    p = fma(r, p, kExp[6]);
    p = fma(r, p, kExp[7]);
    p = fma(r, p, kExp[8]);
    p = fma(r, p, kExp[9]);
    p = fma(r, p, kExp[10]);
    p = fma(r, p, kExp[11]);
    p = fma(r, p, kExp[12]);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// one sphere pair, d = p_owner - p_other: accumulates the sigmoid into phi (if WITH_PHI) and the unit vector into (gx, gy, gz)
template <bool WITH_PHI>
__device__ __forceinline__ void sph_pair(double dx, double dy, double dz, double rsum, double k5, double& phi, double& gx, double& gy, double& gz,
                                         double& dmin)
{
    const double n2 = fma(dx, dx, fma(dy, dy, dz * dz));
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(n2));
    // one third-order step from the 2^-22 seed: y <- y (1 + e / 2 + 3 e^2 / 8), e = 1 - n2 y^2 (error ~ e^3: full double precision;
    // five FP64 instructions where two Newton steps took seven, and this kernel is bound by the FP64 issue rate)
    const double e1 = fma(-(n2 * y), y, 1.0);
    y = fma(y, fma(e1, 0.375, 0.5) * e1, y);  // 1 / |d|  (inf -> NaN for coincident centres, like the reference's division by zero)
    gx = fma(dx, y, gx);
    gy = fma(dy, y, gy);
    gz = fma(dz, y, gz);
    if (WITH_PHI)
    {
        const double n = (n2 == 0.0) ? 0.0 : n2 * y;
        const double gap = n - rsum;  // surface distance of the pair
        dmin = gap < dmin ? gap : dmin;
        const double den = 1.0 + sph_exp(k5 * gap);
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
        const double e = fma(-den, r, 1.0);
        r = fma(r, fma(e, e, e), r);  // r (1 + e + e^2): third order, three instructions instead of two Newton steps' four
        phi += r;
    }
}

#ifndef EXO_SPH_UNROLL
#define EXO_SPH_UNROLL 4
#endif
#ifndef EXO_SPHERES_MAXNREG
#define EXO_SPHERES_MAXNREG 96
#endif
#ifndef EXO_SPHERES_EPC
#define EXO_SPHERES_EPC 4
#endif

// this lane's share of the other group [t0, t1) (stride parts) against its owner sphere: four pairs per iteration -- straight-line
// code whose independent Horner / Newton chains the scheduler interleaves (the FP64 pipe needs ~5 independent instructions in
// flight per warp); the tail (at most three pairs per lane) runs one at a time
template <bool WITH_PHI>
__device__ __forceinline__ void sph_pair_loop(const double2* c2, int j0, int t1, int parts, double pix, double piy, double piz, double ri, double k5,
                                              double& ph, double& gx, double& gy, double& gz, double& dmin)
{
    for (; j0 + (EXO_SPH_UNROLL - 1) * parts < t1; j0 += EXO_SPH_UNROLL * parts)
    {
#pragma unroll
        for (int u = 0; u < EXO_SPH_UNROLL; ++u)
        {
            const int j = j0 + u * parts;
            const double2 a = c2[2 * j], b = c2[2 * j + 1];  // (x, y), (z, radius)
            sph_pair<WITH_PHI>(pix - a.x, piy - a.y, piz - b.x, ri + b.y, k5, ph, gx, gy, gz, dmin);
        }
    }
    for (; j0 < t1; j0 += parts)
    {
        const double2 a = c2[2 * j0], b = c2[2 * j0 + 1];
        sph_pair<WITH_PHI>(pix - a.x, piy - a.y, piz - b.x, ri + b.y, k5, ph, gx, gy, gz, dmin);
    }
}

template <int CS>
__device__ __forceinline__ void emit_spheres_warp(const Op& op, const double* c, const int32_t* tab, int N, int joint_base, double* scr0,
                                               double* warp_smem, int lane, int warp_col0, int n_valid, int64_t e0, const KernelParams& P,
                                               bool want_j)
{
    const int S = op.n_items;
    const int G = tab[0];
    const int32_t* gs = tab + 2;
    const int32_t* sph = gs + G + 1;
    const int32_t* cols = sph + SPH_INTS * S;
    const double k5 = c[0];
    const double* sc = c + 1;  // per sphere: ox oy oz radius
    double* centre = warp_smem;       // [S][4]: x y z radius of the evaluation being processed
    double* rowacc = warp_smem + 4 * S;  // [N]
    for (int ev = 0; ev < n_valid; ++ev)
    {
        const double* scr = scr0 + warp_col0 + ev;  // scratch column of the evaluation being processed
        // ---- sphere centres of this evaluation ---------------------------------------------------------------
        for (int s = lane; s < S; s += 32)
        {
            const int src = sph[SPH_INTS * s];
            double px = sc[4 * s], py = sc[4 * s + 1], pz = sc[4 * s + 2];
            if (src != SRC_ROOT)
            {
                const Fr f = fr_load_scratch<CS>(scr + 12 * src * CS);
                const double tx = f.m0 * px + f.m1 * py + f.m2 * pz + f.x;
                const double ty = f.m3 * px + f.m4 * py + f.m5 * pz + f.y;
                const double tz = f.m6 * px + f.m7 * py + f.m8 * pz + f.z;
                px = tx; py = ty; pz = tz;
            }
            centre[4 * s] = px; centre[4 * s + 1] = py; centre[4 * s + 2] = pz; centre[4 * s + 3] = sc[4 * s + 3];
        }
        __syncwarp();
        const double2* c2 = reinterpret_cast<const double2*>(centre);  // [S][2]: (x, y), (z, radius)
        const int64_t e = e0 + ev;
        int out_idx = 0;
        double dmin = CUDART_INF;  // smallest surface distance over all pairs of this evaluation
        for (int ga = 0; ga < G; ++ga)
        {
            for (int gb = ga + 1; gb < G; ++gb, ++out_idx)
            {
                // Which of the two groups hang below controlled joints?  Only those contribute Jacobian terms.  With one moving
                // group (robot against fixed or trajectory-driven obstacles) a single sweep over the pairs, owned by the moving
                // group's spheres, yields phi and J; with two, a second sweep owned by group B adds its J terms (no sigmoids).
                bool mov_a = false, mov_b = false;
                if (want_j)
                {
                    for (int j = gs[ga] + lane; j < gs[ga + 1]; j += 32) mov_a = mov_a || sph[SPH_INTS * j + 1] >= 0;
                    for (int j = gs[gb] + lane; j < gs[gb + 1]; j += 32) mov_b = mov_b || sph[SPH_INTS * j + 1] >= 0;
                    mov_a = __any_sync(0xffffffffu, mov_a);
                    mov_b = __any_sync(0xffffffffu, mov_b);
                    for (int col = lane; col < N; col += 32) rowacc[col] = 0.0;
                    __syncwarp();
                }
                double phi_part = 0.0;
                const int n_pass = (mov_a && mov_b) ? 2 : 1;
                for (int pass = 0; pass < n_pass; ++pass)
                {
                    const bool own_b = pass == 1 || (mov_b && !mov_a);
                    const bool with_j = own_b ? mov_b : mov_a;
                    const int o0 = own_b ? gs[gb] : gs[ga], o1 = own_b ? gs[gb + 1] : gs[ga + 1];
                    const int t0 = own_b ? gs[ga] : gs[gb], t1 = own_b ? gs[ga + 1] : gs[gb + 1];
                    const int n_own = o1 - o0;
                    // lanes = (owner sphere, share of the other group): parts = how many lanes share one owner sphere
                    int own_p2 = 1;
                    while (own_p2 < n_own && own_p2 < 32) own_p2 <<= 1;
                    const int parts = 32 / own_p2;
                    for (int i0 = o0; i0 < o1; i0 += own_p2)
                    {
                        const int i = i0 + (lane & (own_p2 - 1));
                        const int part = lane / own_p2;
                        const bool live = i < o1;
                        const int ii = live ? i : o1 - 1;
                        const double2 pa = c2[2 * ii], pb = c2[2 * ii + 1];
                        const double pix = pa.x, piy = pa.y, piz = pb.x, ri = pb.y;
                        // d = p_owner - p_other in both sweeps: for owners in group B that is -u_ij, and the sum over the owner's
                        // pairs is -h_j, so both sweeps subtract (w_c x (p_owner - p_c)) . g  (J.row = -sum (J_i - J_j)^T u_ij)
                        double gx = 0.0, gy = 0.0, gz = 0.0, ph = 0.0;
                        if (live)
                        {
                            if (pass == 0)
                                sph_pair_loop<true>(c2, t0 + part, t1, parts, pix, piy, piz, ri, k5, ph, gx, gy, gz, dmin);
                            else
                                sph_pair_loop<false>(c2, t0 + part, t1, parts, pix, piy, piz, ri, k5, ph, gx, gy, gz, dmin);
                        }
                        phi_part += ph;
                        if (!with_j) continue;
                        // Jacobian: every column whose joint carries this lane's sphere gets w_c . ((p_i - p_c) x g).  Eight columns
                        // at a time: each lane evaluates its eight terms, then a butterfly that halves the number of live values
                        // per step (4 + 2 + 1 + 1 + 1 exchanges instead of 8 x 5) leaves column (lane >> 2) in every lane.
                        const int acc_i = live ? sph[SPH_INTS * ii + 1] : -1;
                        for (int col0 = 0; col0 < N; col0 += 8)
                        {
                            double v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                            {
                                v[u] = 0.0;
                                const int col = col0 + u;
                                if (col >= N) continue;
                                const int ent = cols[3 * col];
                                if (ent < 0) continue;
                                const uint32_t mlo = (uint32_t)cols[3 * col + 1], mhi = (uint32_t)cols[3 * col + 2];
                                const bool in = acc_i >= 0 && (acc_i < 32 ? ((mlo >> acc_i) & 1u) : ((mhi >> (acc_i - 32)) & 1u));
                                if (!in) continue;
                                const double* js = scr + (joint_base + 6 * (ent >> 2)) * CS;
                                const double ax = js[0], ay = js[CS], az = js[2 * CS];
                                if (ent & 2)
                                    v[u] = ax * gx + ay * gy + az * gz;
                                else
                                {
                                    const double rx = pix - js[3 * CS], ry = piy - js[4 * CS], rz = piz - js[5 * CS];
                                    v[u] = ax * (ry * gz - rz * gy) + ay * (rz * gx - rx * gz) + az * (rx * gy - ry * gx);
                                }
                            }
                            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                const double send = h16 ? v[u] : v[u + 4], keep = h16 ? v[u + 4] : v[u];
                                v[u] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                            }
#pragma unroll
                            for (int u = 0; u < 2; ++u)
                            {
                                const double send = h8 ? v[u] : v[u + 2], keep = h8 ? v[u + 2] : v[u];
                                v[u] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                            }
                            {
                                const double send = h4 ? v[0] : v[1], keep = h4 ? v[1] : v[0];
                                v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                            }
                            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
                            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
                            const int mycol = col0 + (lane >> 2);  // bit 2 picked between 2 values, bit 3 between 4, bit 4 between 8
                            if ((lane & 3) == 0 && mycol < N) rowacc[mycol] -= v[0];
                        }
                    }
                }
                phi_part = warp_sum(phi_part);
                if (lane == 0) P.phi[e * P.phi_se + (int64_t)(op.phi_off + out_idx) * P.phi_sk] = phi_part;
                if (want_j)
                {
                    __syncwarp();
                    double* row = P.jac + e * P.j_se + (int64_t)(op.jrow_off + out_idx) * P.j_sr;
                    for (int col = lane; col < N; col += 32) row[(int64_t)col * P.j_sc] = rowacc[col];
                }
                __syncwarp();
            }
        }
        if (P.min_dist)
        {
            // new output (not in the reference): what the multi-GPU path all-reduces with min
            for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
            if (lane == 0) P.min_dist[e] = (op.flags & 1) ? dmin : fmin(P.min_dist[e], dmin);
        }
        __syncwarp();  // the centre table is rewritten for the next evaluation
    }
}

// TPB threads per CTA; W evaluations per warp tile (W < 32: lanes l and l + W work on the same evaluation -- redundant
// forward kinematics in exchange for a W/32 scratch footprint, used when SphereCollision needs many resident warps).
template <int TPB, int W>
__global__ void __maxnreg__(168) exo_eval_kernel(const KernelParams P)
{
    // scratch columns per CTA; full-width tiles get one padding column so that a warp can also read one COLUMN (consecutive
    // words of one evaluation, stride CS doubles) without bank conflicts -- the split SphereCollision launch parks rows that way
    constexpr int CS = TPB / 32 * W + (W == 32 ? 1 : 0);
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    // ---- stage the program (topology, axes, offsets, tables) once per CTA ---------------------------------
    {
        const int n_words = P.program_bytes >> 2;
        const int32_t* src = reinterpret_cast<const int32_t*>(P.program);
        int32_t* dst = reinterpret_cast<int32_t*>(smem);
        for (int i = tid; i < n_words; i += TPB) dst[i] = src[i];  // CTA-wide copy
    }
    __syncthreads();
    const ProgramHeader* H = reinterpret_cast<const ProgramHeader*>(smem);
    const Op* ops = reinterpret_cast<const Op*>(smem + sizeof(ProgramHeader));
    const int32_t* ints = reinterpret_cast<const int32_t*>(ops + H->n_ops);
    const double* dbls = reinterpret_cast<const double*>(smem + P.dbl_offset);
    const int lane = tid & 31;
    const int col = (tid >> 5) * W + (lane & (W - 1));                        // this lane's scratch column
    double* scr = reinterpret_cast<double*>(smem + P.scratch_offset) + col;  // word k at scr[k * CS]
    const int N = H->n_ctrl;
    const int joint_base = 12 * H->n_frame_slots;
    const int pos_base = joint_base;  // two-pass programs park 3 doubles per requested frame where the joint slots would be
    double* qs = scr + (size_t)H->scratch_doubles * CS;
    // per-warp area behind the per-thread scratch: sphere centre table (4 doubles per sphere) + Jacobian row accumulators
    double* warp_smem = reinterpret_cast<double*>(smem + P.scratch_offset) + ((((size_t)(H->scratch_doubles + N) * CS) + 1) & ~(size_t)1) +
                        (size_t)(tid >> 5) * (4 * H->max_spheres + N + (N & 1));  // even number of doubles: 16-byte aligned tables
    const bool want_j = P.jac != nullptr;

    // The loop is warp-uniform (SphereCollision is evaluated by the whole warp for one lane's evaluation at a time):
    // lanes past the end of the batch redo the last evaluation and write the same values to the same addresses.
    for (int64_t e0 = ((int64_t)blockIdx.x * (TPB / 32) + (tid >> 5)) * W; e0 < P.n_evals; e0 += (int64_t)gridDim.x * (TPB / 32) * W)
    {
        const int n_valid = (int)((P.n_evals - e0) < W ? (P.n_evals - e0) : W);  // evaluations of this warp's tile
        const int slot = lane & (W - 1);
        const int64_t e = e0 + (slot < n_valid ? slot : n_valid - 1);
        // ---- stage this evaluation's state (N independent loads in flight) --------------------------------
        {
            const double* xe = P.x + e * P.x_se;
            for (int i = 0; i < N; ++i) qs[i * CS] = xe[i * P.x_sk];
        }
        double* phi_e = P.phi + e * P.phi_se;
        double* jac_e = want_j ? P.jac + e * P.j_se : nullptr;
        Fr cur = fr_identity();
        int sph_ord = 0;  // SphereCollision ops seen so far in this evaluation (indexes the dump of the split launch)
        // two-pass programs: pass 0 composes the tree and emits Phi (parking each requested frame's position), pass 1
        // composes it again and emits every joint's Jacobian columns the moment its axis and origin are known
        constexpr bool TP = W == 32;  // narrow tiles serve SphereCollision programs, which are never two-pass: compile it out
        const bool two_pass = TP && H->two_pass;
        const int n_pass = (two_pass && want_j) ? 2 : 1;
        for (int pass = 0; pass < n_pass; ++pass)
        for (int pc = 0;; ++pc)
        {
            const Op op = ops[pc];
            if (op.kind == OP_END) break;
            if (pass == 1 && op.kind >= OP_EMIT_FRAME) continue;
            const double* c = dbls + op.cst;
            if (op.kind == OP_JOINT || op.kind == OP_TABLE)
            {
                if (op.src == SRC_ROOT)
                    cur = fr_identity();
                else if (op.src >= 0)
                    cur = fr_load_scratch<CS>(scr + 12 * op.src * CS);
                const int pre = op.flags & F_PRE_MASK;
                if (pre == F_PRE_TRANS)
                    fr_translate(cur, c[9], c[10], c[11]);
                else if (pre == F_PRE_FULL)
                    cur = fr_mul(cur, fr_load_const(c));
                if (op.kind == OP_TABLE)
                {
                    const int t = (int)((P.e_base + e) % P.T);
                    const double* g = P.generated + ((size_t)t * H->n_generated + op.qidx) * 12;
                    Fr gf;
                    gf.m0 = g[0]; gf.m1 = g[1]; gf.m2 = g[2]; gf.m3 = g[3]; gf.m4 = g[4]; gf.m5 = g[5];
                    gf.m6 = g[6]; gf.m7 = g[7]; gf.m8 = g[8]; gf.x = g[9]; gf.y = g[10]; gf.z = g[11];
                    cur = fr_mul(cur, gf);
                }
                else
                {
                    // z-canonical joint (compiler.cpp): the joint moves about / along the z axis of the frame in front of it,
                    // so its world axis is the z column and its origin the translation, whatever the URDF axis was
                    double q = c[CJ_QOFF];
                    if (op.qidx >= 0) q = fma(c[CJ_QMULT], qs[op.qidx * CS], q);
                    const double wx = cur.m2, wy = cur.m5, wz = cur.m8;
                    if (op.flags & F_PRISMATIC)
                    {
                        cur.x = fma(wx, q, cur.x);
                        cur.y = fma(wy, q, cur.y);
                        cur.z = fma(wz, q, cur.z);
                    }
                    else
                    {
                        double s, cs;
                        if (fabs(q) <= 1.0e5)
                            fast_sincos(q, s, cs);
                        else
                            sincos(q, &s, &cs);  // Payne-Hanek territory: never for real joint angles
                        // M * Rz: col0' = c col0 + s col1, col1' = -s col0 + c col1
                        const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                        cur.m0 = fma(cs, a0, s * cur.m1); cur.m1 = fma(cs, cur.m1, -(s * a0));
                        cur.m3 = fma(cs, a3, s * cur.m4); cur.m4 = fma(cs, cur.m4, -(s * a3));
                        cur.m6 = fma(cs, a6, s * cur.m7); cur.m7 = fma(cs, cur.m7, -(s * a6));
                    }
                    if (two_pass)
                    {
                        if (pass == 1)
                        {
                            // ComputeJ (kinematic_tree.cpp:919-945) for every requested frame below this joint
                            const int32_t* tp = ints + op.tab;
                            const int n_ent = tp[0];
                            for (int k = 0; k < n_ent; ++k)
                            {
                                const int32_t* ent = tp + 1 + TP_INTS * k;
                                const double* ps = scr + (pos_base + 3 * ent[0]) * CS;
                                double v0, v1, v2, w0 = 0.0, w1 = 0.0, w2 = 0.0;
                                if (ent[4] & TP_F_PRISMATIC)
                                {
                                    v0 = wx; v1 = wy; v2 = wz;
                                }
                                else
                                {
                                    const double dx = ps[0] - cur.x, dy = ps[CS] - cur.y, dz = ps[2 * CS] - cur.z;
                                    v0 = wy * dz - wz * dy;
                                    v1 = wz * dx - wx * dz;
                                    v2 = wx * dy - wy * dx;
                                    w0 = wx; w1 = wy; w2 = wz;
                                }
                                if (ent[4] & TP_F_BCONST)
                                {
                                    const double* b = dbls + ent[5];
                                    const double t0 = b[0] * v0 + b[1] * v1 + b[2] * v2;
                                    const double t1 = b[3] * v0 + b[4] * v1 + b[5] * v2;
                                    const double t2 = b[6] * v0 + b[7] * v1 + b[8] * v2;
                                    v0 = t0; v1 = t1; v2 = t2;
                                    const double u0 = b[0] * w0 + b[1] * w1 + b[2] * w2;
                                    const double u1 = b[3] * w0 + b[4] * w1 + b[5] * w2;
                                    const double u2 = b[6] * w0 + b[7] * w1 + b[8] * w2;
                                    w0 = u0; w1 = u1; w2 = u2;
                                }
                                const int64_t sr = P.j_sr;
                                // EffAxisAlignment parked its vector where the position would be; CenterOfMass's mass share
                                // sits behind the frame's constants
                                const double ux = ent[3] >= ROWS_V_ACC ? dbls[ent[5] - CE_B_M + CE_SIZE] : ps[0];
                                store_rows(ent[3], jac_e + (int64_t)ent[2] * sr + (int64_t)ent[1] * P.j_sc, sr, v0, v1, v2, w0, w1, w2, ux, ps[CS], ps[2 * CS]);
                            }
                        }
                    }
                    else if (op.jslot >= 0)
                    {
                        double* js = scr + (joint_base + 6 * op.jslot) * CS;
                        js[0] = wx; js[CS] = wy; js[2 * CS] = wz;
                        js[3 * CS] = cur.x; js[4 * CS] = cur.y; js[5 * CS] = cur.z;
                    }
                }
                if (op.dst >= 0) fr_store_scratch<CS>(scr + 12 * op.dst * CS, cur);
            }
            else if (op.kind == OP_EMIT_FRAME)
            {
                // temp_A = frame_A * offA ; temp_B = frame_B * offB ; Phi = temp_B^-1 * temp_A  (kinematic_tree.cpp:829-835)
                Fr A;
                if (op.src == SRC_CUR) A = cur;
                else if (op.src == SRC_ROOT) A = fr_identity();
                else A = fr_load_scratch<CS>(scr + 12 * op.src * CS);
                const int offa = op.flags & F_OFFA_MASK;
                if (offa == F_OFFA_TRANS) fr_translate(A, c[CE_OFFA_P], c[CE_OFFA_P + 1], c[CE_OFFA_P + 2]);
                else if (offa == F_OFFA_FULL) A = fr_mul(A, fr_load_const(c + CE_OFFA_M));
                const int bk = op.flags & F_B_MASK;
                Fr Binv, Phi;
                if (bk == F_B_IDENTITY)
                    Phi = A;
                else
                {
                    if (bk == F_B_CONST)
                        Binv = fr_load_const(c + CE_B_M);
                    else
                    {
                        Fr B;
                        if (op.src_b == SRC_CUR) B = cur;
                        else B = fr_load_scratch<CS>(scr + 12 * op.src_b * CS);
                        B = fr_mul(B, fr_load_const(c + CE_B_M));
                        Binv = fr_inverse(B);
                    }
                    Phi = fr_mul(Binv, A);
                }
                const int row_mode = (op.flags & F_ROWS_MASK) >> F_ROWS_SHIFT;
                double ux = 0.0, uy = 0.0, uz = 0.0;  // per-frame vector / weight some row modes need (store_rows)
                if (row_mode == ROWS_WDOT)
                {
                    // EffAxisAlignment::Update (eff_axis_alignment.cpp:116-145): phi = direction . (p_{i+n} - p_i) - 1
                    const double* k = c + CE_SIZE;  // a (3), direction (3)
                    const double u0 = Phi.m0 * k[0] + Phi.m1 * k[1] + Phi.m2 * k[2];
                    const double u1 = Phi.m3 * k[0] + Phi.m4 * k[1] + Phi.m5 * k[2];
                    const double u2 = Phi.m6 * k[0] + Phi.m7 * k[1] + Phi.m8 * k[2];
                    phi_e[(int64_t)op.phi_off * P.phi_sk] = (u0 * k[3] + u1 * k[4] + u2 * k[5]) - 1.0;
                    ux = u1 * k[5] - u2 * k[4];
                    uy = u2 * k[3] - u0 * k[5];
                    uz = u0 * k[4] - u1 * k[3];
                }
                else if (row_mode >= ROWS_V_ACC)
                {
                    // CenterOfMass::Update (center_of_mass.cpp:37-54): com += p_i mass_i / M on rows a fill op zeroed
                    ux = c[CE_SIZE];
                    double* ph = phi_e + (int64_t)op.phi_off * P.phi_sk;
                    ph[0] += ux * Phi.x;
                    ph[P.phi_sk] += ux * Phi.y;
                    if (row_mode == ROWS_V_ACC) ph[2 * P.phi_sk] += ux * Phi.z;
                }
                else
                    write_phi(op.map_kind, op.rot_type, Phi, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk);
                if (want_j && two_pass)
                {
                    // park temp_A.p for pass 1 (EffAxisAlignment: the vector its row is dotted with instead) and write the
                    // structural zeros (columns no joint below this frame owns)
                    double* ps = scr + (pos_base + 3 * op.n_items) * CS;
                    if (row_mode == ROWS_WDOT)
                    {
                        ps[0] = ux; ps[CS] = uy; ps[2 * CS] = uz;
                    }
                    else
                    {
                        ps[0] = A.x; ps[CS] = A.y; ps[2 * CS] = A.z;
                    }
                    const int32_t* zl = ints + op.pad;
                    const int n_zero = zl[0];
                    const int64_t sr = P.j_sr, sc = P.j_sc;
                    double* jr = jac_e + (int64_t)op.jrow_off * sr;
                    const int n_rows = op.n_rows;
                    for (int k = 1; k <= n_zero; ++k)
                    {
                        double* jc = jr + (int64_t)zl[k] * sc;
                        for (int r = 0; r < n_rows; ++r, jc += sr) EXO_ST(jc, 0.0);
                    }
                }
                else if (want_j)
                {
                    // ComputeJ (kinematic_tree.cpp:919-945): column c = temp_B.M^-1 * [ w x (temp_A.p - p_joint) ; w ]
                    const int32_t* tab = ints + op.tab;
                    double* jr = jac_e + (int64_t)op.jrow_off * P.j_sr;
                    const int64_t sr = P.j_sr;
                    if (row_mode == ROWS_DIST)
                    {
                        const double nrm = norm3_scaled(Phi.x, Phi.y, Phi.z);
                        ux = Phi.x / nrm; uy = Phi.y / nrm; uz = Phi.z / nrm;
                    }
                    for (int col = 0; col < N; ++col)
                    {
                        const int ent = tab[col];
                        if (ent < 0 && row_mode >= ROWS_V_ACC) continue;  // accumulating rows: nothing to add
                        double v0 = 0.0, v1 = 0.0, v2 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
                        if (ent >= 0)
                        {
                            const double* js = scr + (joint_base + 6 * (ent >> 2)) * CS;
                            const double ax = js[0], ay = js[CS], az = js[2 * CS];
                            if (ent & 2)
                            {
                                v0 = ax; v1 = ay; v2 = az;
                            }
                            else
                            {
                                const double dx = A.x - js[3 * CS], dy = A.y - js[4 * CS], dz = A.z - js[5 * CS];
                                v0 = ay * dz - az * dy;
                                v1 = az * dx - ax * dz;
                                v2 = ax * dy - ay * dx;
                                w0 = ax; w1 = ay; w2 = az;
                            }
                            if (bk != F_B_IDENTITY)
                            {
                                const double t0 = Binv.m0 * v0 + Binv.m1 * v1 + Binv.m2 * v2;
                                const double t1 = Binv.m3 * v0 + Binv.m4 * v1 + Binv.m5 * v2;
                                const double t2 = Binv.m6 * v0 + Binv.m7 * v1 + Binv.m8 * v2;
                                v0 = t0; v1 = t1; v2 = t2;
                                const double u0 = Binv.m0 * w0 + Binv.m1 * w1 + Binv.m2 * w2;
                                const double u1 = Binv.m3 * w0 + Binv.m4 * w1 + Binv.m5 * w2;
                                const double u2 = Binv.m6 * w0 + Binv.m7 * w1 + Binv.m8 * w2;
                                w0 = u0; w1 = u1; w2 = u2;
                            }
                            if (ent & 1)
                            {
                                // frame-B chain: 0 - column (kinematic_tree.cpp:941)
                                v0 = 0.0 - v0; v1 = 0.0 - v1; v2 = 0.0 - v2;
                                w0 = 0.0 - w0; w1 = 0.0 - w1; w2 = 0.0 - w2;
                            }
                        }
                        store_rows(row_mode, jr + (int64_t)col * P.j_sc, sr, v0, v1, v2, w0, w1, w2, ux, uy, uz);
                    }
                }
            }
            else if (op.kind == OP_EMIT_JLIMIT)
            {
                emit_joint_limit(c, N, qs, CS, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, want_j ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            }
            else if (op.kind == OP_EMIT_JPOSE)
            {
                emit_joint_pose(c, ints + op.tab, op.n_items, N, qs, CS, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, want_j ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            }
            else if (op.kind == OP_EMIT_FILL)
            {
                emit_fill(c, op.n_items, op.n_rows, N, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, want_j ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            }
            else if (op.kind == OP_EMIT_SPHERES && P.sph_dump)
            {
                // split launch: park every evaluation's slots for exo_spheres_kernel, one contiguous row each.  The warp transposes
                // its tile on the way out: for one evaluation at a time, lane k copies words k, k + 32, ... (coalesced 256-byte stores;
                // the shared-memory side is conflict-free thanks to the padding column)
                __syncwarp();
                {
                    const int words = H->scratch_doubles;
                    const double* tile = scr - col + (tid >> 5) * W;  // column of the tile's first evaluation
                    for (int ev = 0; ev < n_valid; ++ev)
                    {
                        double* d = P.sph_dump + (int64_t)sph_ord * P.sph_block + (e0 + ev) * P.sph_row;
                        for (int k = lane; k < words; k += 32) d[k] = tile[ev + k * CS];
                    }
                }
                __syncwarp();
                ++sph_ord;
            }
            else if (op.kind == OP_EMIT_SPHERES)
            {
                __syncwarp();  // every lane's frame / joint slots are complete
                emit_spheres_warp<CS>(op, c, ints + op.tab, N, joint_base, scr - col, warp_smem, lane, (tid >> 5) * W,
                                       n_valid, e0, P, want_j);
                __syncwarp();  // slots may be overwritten by the ops that follow
            }
        }
    }
}

// SphereCollision of split programs: one warp per evaluation, WPC independent warps per CTA.  A warp copies its evaluation's row of
// parked slots into its own shared-memory buffer with 16-byte cp.async while it is still busy with the previous row (two buffers),
// so no warp ever waits for another; the pair loops are the interpreter's own code with a scratch stride of one column.  Without
// the tree interpreter around them they need 96 registers instead of 166.
template <int WPC>
__global__ void __maxnreg__(EXO_SPHERES_MAXNREG) exo_spheres_kernel(const KernelParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int sph_pc[SPH_SPLIT_MAX_OPS];
    const int tid = threadIdx.x;
    {
        const int n_words = P.program_bytes >> 2;
        const int32_t* src = reinterpret_cast<const int32_t*>(P.program);
        int32_t* dst = reinterpret_cast<int32_t*>(smem);
        for (int i = tid; i < n_words; i += WPC * 32) dst[i] = src[i];
    }
    __syncthreads();
    const ProgramHeader* H = reinterpret_cast<const ProgramHeader*>(smem);
    const Op* ops = reinterpret_cast<const Op*>(smem + sizeof(ProgramHeader));
    const int32_t* ints = reinterpret_cast<const int32_t*>(ops + H->n_ops);
    const double* dbls = reinterpret_cast<const double*>(smem + P.dbl_offset);
    if (tid == 0)
    {
        int k = 0;
        for (int pc = 0; ops[pc].kind != OP_END; ++pc)
            if (ops[pc].kind == OP_EMIT_SPHERES && k < SPH_SPLIT_MAX_OPS) sph_pc[k++] = pc;
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int N = H->n_ctrl;
    const int row = P.sph_row, n_ops = P.sph_n_ops;
    const int joint_base = 12 * H->n_frame_slots;
    const int per_warp = 2 * row + 4 * H->max_spheres + N + (N & 1);  // two row buffers, centre table, Jacobian row accumulators
    double* stage = reinterpret_cast<double*>(smem + P.scratch_offset) + (size_t)warp * per_warp;
    double* warp_smem = stage + 2 * row;
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const bool want_j = P.jac != nullptr;
    const int64_t stride = (int64_t)gridDim.x * WPC;

    auto fetch = [&](int64_t e, int ord, int buf) {
        const uint32_t dst = stage_s + (uint32_t)(buf * row) * 8u;
        if (P.sph_tiled)
        {
            // rows written by the generated park kernel: word k of evaluation e at [e / 32][k][e % 32]
            const char* src = reinterpret_cast<const char*>(P.sph_dump + (int64_t)ord * P.sph_block + (e & ~(int64_t)31) * row + (e & 31));
            for (int i = lane; i < row; i += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * i), "l"(src + 256 * i) : "memory");
        }
        else
        {
            const char* src = reinterpret_cast<const char*>(P.sph_dump + (int64_t)ord * P.sph_block + e * row);
            for (int i = lane; i < (row >> 1); i += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int64_t e = (int64_t)blockIdx.x * WPC + warp;
    int ord = 0, buf = 0;
    if (e < P.n_evals) fetch(e, 0, 0);
    while (e < P.n_evals)
    {
        int64_t e_next = e;
        int ord_next = ord + 1;
        if (ord_next == n_ops)
        {
            ord_next = 0;
            e_next = e + stride;
        }
        if (e_next < P.n_evals)
        {
            fetch(e_next, ord_next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();  // every lane's copies of this row have landed
        const Op op = ops[sph_pc[ord]];
        emit_spheres_warp<1>(op, dbls + op.cst, ints + op.tab, N, joint_base, stage + buf * row, warp_smem, lane, 0, 1, e, P, want_j);
        __syncwarp();  // the buffer is refilled two rows from now
        e = e_next;
        ord = ord_next;
        buf ^= 1;
    }
}
}  // namespace

template <int TPB, int W>
cudaError_t try_config(size_t prog, size_t per_column, size_t per_warp, int max_smem, int& ctas_per_sm, size_t& smem_bytes)
{
    // CS of the kernel (padding column); the per-warp tables behind the scratch start on a 16-byte boundary
    smem_bytes = prog + ((per_column * (TPB / 32 * W + (W == 32 ? 1 : 0)) + 15) & ~(size_t)15) + per_warp * (TPB / 32);
    ctas_per_sm = 0;
    if (smem_bytes > (size_t)max_smem) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(exo_eval_kernel<TPB, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, exo_eval_kernel<TPB, W>, TPB, smem_bytes);
}

namespace
{
struct EvalConfig
{
    int tpb, w;
};
// without SphereCollision: one lane per evaluation, pick the CTA size that keeps most lanes resident;
// with SphereCollision (one warp per evaluation in the pair loops): pick the configuration that keeps most WARPS resident.
const EvalConfig kPlain[] = {{128, 32}, {64, 32}, {32, 32}};
const EvalConfig kSpheres[] = {{64, 8}, {64, 16}, {128, 8}, {64, 32}, {32, 32}};

cudaError_t try_any(const EvalConfig& c, size_t prog, size_t per_column, size_t per_warp, int max_smem, int& ctas, size_t& smem_bytes)
{
    if (c.tpb == 128 && c.w == 32) return try_config<128, 32>(prog, per_column, per_warp, max_smem, ctas, smem_bytes);
    if (c.tpb == 64 && c.w == 32) return try_config<64, 32>(prog, per_column, per_warp, max_smem, ctas, smem_bytes);
    if (c.tpb == 32 && c.w == 32) return try_config<32, 32>(prog, per_column, per_warp, max_smem, ctas, smem_bytes);
    if (c.tpb == 64 && c.w == 8) return try_config<64, 8>(prog, per_column, per_warp, max_smem, ctas, smem_bytes);
    if (c.tpb == 64 && c.w == 16) return try_config<64, 16>(prog, per_column, per_warp, max_smem, ctas, smem_bytes);
    if (c.tpb == 128 && c.w == 8) return try_config<128, 8>(prog, per_column, per_warp, max_smem, ctas, smem_bytes);
    return cudaErrorInvalidConfiguration;
}
}  // namespace

cudaError_t plan_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, bool split_spheres, LaunchConfig& cfg, KernelParams& params)
{
    int dev = 0, sms = 0, max_smem = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (err != cudaSuccess) return err;
    const size_t prog = ((size_t)program_bytes + 15) & ~(size_t)15;
    const size_t per_column = (size_t)(hdr.scratch_doubles + hdr.n_ctrl) * sizeof(double);
    const bool spheres = hdr.max_spheres > 0 && !split_spheres;
    const size_t per_warp = hdr.max_spheres > 0 ? (size_t)(4 * hdr.max_spheres + hdr.n_ctrl + (hdr.n_ctrl & 1)) * sizeof(double) : 0;
    params.scratch_offset = (int32_t)prog;
    const EvalConfig* cands = spheres ? kSpheres : kPlain;
    const int n_cands = spheres ? (int)(sizeof(kSpheres) / sizeof(kSpheres[0])) : (int)(sizeof(kPlain) / sizeof(kPlain[0]));
    int best_score = 0, best_ctas = 0;
    EvalConfig best = {0, 0};
    size_t best_smem = 0;
    for (int k = 0; k < n_cands; ++k)
    {
        int ctas = 0;
        size_t smem_bytes = 0;
        err = try_any(cands[k], prog, per_column, spheres ? per_warp : 0, max_smem, ctas, smem_bytes);
        if (err != cudaSuccess) return err;
        // spheres: resident warps first, wider tiles break ties; plain: resident lanes
        const int score = spheres ? ctas * (cands[k].tpb / 32) * 64 + cands[k].w : ctas * cands[k].tpb;
        if (ctas > 0 && score > best_score)
        {
            best_score = score;
            best = cands[k];
            best_ctas = ctas;
            best_smem = smem_bytes;
        }
    }
    if (best_score == 0) return cudaErrorInvalidConfiguration;
    cfg.threads = best.tpb;
    cfg.tile = best.w;
    cfg.smem_bytes = best_smem;
    const int64_t per_cta = (int64_t)(best.tpb / 32) * best.w;
    const int64_t want = (n_evals + per_cta - 1) / per_cta;
    const int64_t resident = (int64_t)sms * best_ctas;
    cfg.blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    cfg.sph_epc = 0;
    if (hdr.max_spheres > 0 && split_spheres)
    {
        constexpr int EPC = EXO_SPHERES_EPC;
        const size_t row = (size_t)((hdr.scratch_doubles + 1) & ~1);
        const size_t smem_bytes = prog + (2 * row * sizeof(double) + per_warp) * EPC;
        if (smem_bytes + 1024 > (size_t)max_smem) return cudaErrorInvalidConfiguration;  // the kernel has a little static shared memory too
        err = cudaFuncSetAttribute(exo_spheres_kernel<EPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (err != cudaSuccess) return err;
        int ctas = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, exo_spheres_kernel<EPC>, EPC * 32, smem_bytes);
        if (err != cudaSuccess) return err;
        if (ctas < 1) return cudaErrorInvalidConfiguration;
        cfg.sph_epc = EPC;
        cfg.sph_blocks = sms * ctas;
        cfg.sph_smem_bytes = smem_bytes;
    }
    return cudaSuccess;
}

cudaError_t launch_spheres(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream)
{
    constexpr int EPC = EXO_SPHERES_EPC;
    if (cfg.sph_epc != EPC) return cudaErrorInvalidConfiguration;
    const int64_t want = (params.n_evals + EPC - 1) / EPC;
    const int blocks = (int)(want < cfg.sph_blocks ? (want > 0 ? want : 1) : cfg.sph_blocks);
    exo_spheres_kernel<EPC><<<blocks, EPC * 32, cfg.sph_smem_bytes, stream>>>(params);
    return cudaGetLastError();
}

cudaError_t launch_eval(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream)
{
#define EXO_LAUNCH(T, WW) exo_eval_kernel<T, WW><<<cfg.blocks, T, cfg.smem_bytes, stream>>>(params)
    if (cfg.threads == 128 && cfg.tile == 32) EXO_LAUNCH(128, 32);
    else if (cfg.threads == 64 && cfg.tile == 32) EXO_LAUNCH(64, 32);
    else if (cfg.threads == 32 && cfg.tile == 32) EXO_LAUNCH(32, 32);
    else if (cfg.threads == 64 && cfg.tile == 8) EXO_LAUNCH(64, 8);
    else if (cfg.threads == 64 && cfg.tile == 16) EXO_LAUNCH(64, 16);
    else if (cfg.threads == 128 && cfg.tile == 8) EXO_LAUNCH(128, 8);
    else return cudaErrorInvalidConfiguration;
#undef EXO_LAUNCH
    return cudaGetLastError();
}

}  // namespace exo
