// chain_kernel.cu -- register-resident kernel for serial-chain programs (sm_100a).
//
// Same tree program as the interpreter in kernels.cu, but for the common case of a serial manipulator (every joint op
// feeds the next, e.g. the LWR): the op walk is unrolled at compile time over the joint ordinal (template NJ), so the
// running frame, every joint's sin/cos and every joint's world axis + origin (what the Jacobian columns need) live in
// registers addressed statically -- no scratch slots, no register moves between interpreter cases.  All NJ sincos are
// issued first (independent, so the FP64 pipe sees NJ-way ILP), then the chain is composed, and each requested frame
// is emitted at its position in the chain.  Per evaluation the only memory traffic is the coalesced read of q and the
// coalesced write of Phi and J: the kernel is bound by HBM, not by FP64.
//
// Reference path replaced: KinematicTree::Update/UpdateTree/FK/ComputeJ (exotica_core/src/kinematic_tree.cpp:678-737,
// 829-867,919-945) + EffFrame/EffPosition/JointLimit::Update (exotations/exotica_core_task_maps/src/*.cpp).
#include "device_math.cuh"
#include "kernels.cuh"

namespace exo
{
namespace
{
using namespace dev;
constexpr int CTPB = 128;

// Emit one requested frame with NA joints of the chain available in registers.
template <int NA, int NJ>
__device__ __forceinline__ void chain_emit_frame(const Op& op, const double* c, const int32_t* tab, int N, int n_joint_ops, const Fr& cur,
                                                 const double (&W)[NJ][3], const double (&PJ)[NJ][3], double* phi_e, double* jac_e,
                                                 const KernelParams& P)
{
    Fr A = (op.src == SRC_CUR) ? cur : fr_identity();
    const int offa = op.flags & F_OFFA_MASK;
    if (offa == F_OFFA_TRANS)
        fr_translate(A, c[CE_OFFA_P], c[CE_OFFA_P + 1], c[CE_OFFA_P + 2]);
    else if (offa == F_OFFA_FULL)
        A = fr_mul(A, fr_load_const(c + CE_OFFA_M));
    const bool b_const = (op.flags & F_B_MASK) == F_B_CONST;
    Fr Binv = fr_identity();
    Fr Phi = A;
    if (b_const)
    {
        Binv = fr_load_const(c + CE_B_M);
        Phi = fr_mul(Binv, A);
    }
    write_phi(op.map_kind, op.rot_type, Phi, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk);
    if (jac_e == nullptr) return;
    double* jr = jac_e + (int64_t)op.jrow_off * P.j_sr;
    const int64_t sr = P.j_sr;
    const bool three = op.n_rows == 3;
    const int32_t* otab = tab + N;  // by joint ordinal
#pragma unroll
    for (int j = 0; j < NA; ++j)
    {
        const int ent = otab[j];
        if (ent < 0) continue;
        double v0, v1, v2, w0 = 0.0, w1 = 0.0, w2 = 0.0;
        if (ent & 2)
        {
            v0 = W[j][0]; v1 = W[j][1]; v2 = W[j][2];
        }
        else
        {
            const double dx = A.x - PJ[j][0], dy = A.y - PJ[j][1], dz = A.z - PJ[j][2];
            v0 = W[j][1] * dz - W[j][2] * dy;
            v1 = W[j][2] * dx - W[j][0] * dz;
            v2 = W[j][0] * dy - W[j][1] * dx;
            w0 = W[j][0]; w1 = W[j][1]; w2 = W[j][2];
        }
        if (b_const)
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
            v0 = 0.0 - v0; v1 = 0.0 - v1; v2 = 0.0 - v2;
            w0 = 0.0 - w0; w1 = 0.0 - w1; w2 = 0.0 - w2;
        }
        double* jc = jr + (int64_t)(ent >> 2) * P.j_sc;
        jc[0] = v0; jc[sr] = v1; jc[2 * sr] = v2;
        if (!three)
        {
            jc[3 * sr] = w0; jc[4 * sr] = w1; jc[5 * sr] = w2;
        }
    }
    // columns without a contributor (joints that are not ancestors of the frame) are written as zeros
    for (int col = 0; col < N; ++col)
    {
        if (tab[col] >= 0) continue;
        double* jc = jr + (int64_t)col * P.j_sc;
        jc[0] = 0.0; jc[sr] = 0.0; jc[2 * sr] = 0.0;
        if (!three)
        {
            jc[3 * sr] = 0.0; jc[4 * sr] = 0.0; jc[5 * sr] = 0.0;
        }
    }
}

// One joint of the chain (ordinal I): cur <- cur * pre * J(u, q); records the joint's world axis and origin.
template <int NJ>
__device__ __forceinline__ void chain_joint(const Op& op, const double* c, Fr& cur, double sn_i, double cs_i, double (&W)[3], double (&PJ)[3])
{
    const int pre = op.flags & F_PRE_MASK;
    if (pre == F_PRE_TRANS)
        fr_translate(cur, c[9], c[10], c[11]);
    else if (pre == F_PRE_FULL)
        cur = fr_mul(cur, fr_load_const(c));
    const int ax = op.flags & F_AXIS_MASK;
    double wx, wy, wz;
    if (ax == F_AXIS_GENERAL)
    {
        const double ux = c[CJ_AXIS], uy = c[CJ_AXIS + 1], uz = c[CJ_AXIS + 2];
        if (!(op.flags & F_PRISMATIC))
        {
            const double s = sn_i, co = cs_i, vt = 1.0 - co;
            Fr r;
            r.m0 = co + vt * ux * ux;
            r.m1 = -uz * s + vt * ux * uy;
            r.m2 = uy * s + vt * ux * uz;
            r.m3 = uz * s + vt * ux * uy;
            r.m4 = co + vt * uy * uy;
            r.m5 = -ux * s + vt * uy * uz;
            r.m6 = -uy * s + vt * ux * uz;
            r.m7 = ux * s + vt * uy * uz;
            r.m8 = co + vt * uz * uz;
            r.x = r.y = r.z = 0.0;
            const double px = cur.x, py = cur.y, pz = cur.z;
            cur = fr_mul(cur, r);
            cur.x = px; cur.y = py; cur.z = pz;
        }
        wx = cur.m0 * ux + cur.m1 * uy + cur.m2 * uz;
        wy = cur.m3 * ux + cur.m4 * uy + cur.m5 * uz;
        wz = cur.m6 * ux + cur.m7 * uy + cur.m8 * uz;
    }
    else
    {
        const bool neg = (op.flags & F_AXIS_NEG) != 0;
        const double sg = neg ? -1.0 : 1.0;
        if (!(op.flags & F_PRISMATIC))
        {
            const double s = neg ? -sn_i : sn_i, co = cs_i;
            if (ax == F_AXIS_X)
            {
                const double a1 = cur.m1, a4 = cur.m4, a7 = cur.m7;
                cur.m1 = co * a1 + s * cur.m2; cur.m2 = co * cur.m2 - s * a1;
                cur.m4 = co * a4 + s * cur.m5; cur.m5 = co * cur.m5 - s * a4;
                cur.m7 = co * a7 + s * cur.m8; cur.m8 = co * cur.m8 - s * a7;
            }
            else if (ax == F_AXIS_Y)
            {
                const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                cur.m0 = co * a0 - s * cur.m2; cur.m2 = s * a0 + co * cur.m2;
                cur.m3 = co * a3 - s * cur.m5; cur.m5 = s * a3 + co * cur.m5;
                cur.m6 = co * a6 - s * cur.m8; cur.m8 = s * a6 + co * cur.m8;
            }
            else
            {
                const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                cur.m0 = co * a0 + s * cur.m1; cur.m1 = co * cur.m1 - s * a0;
                cur.m3 = co * a3 + s * cur.m4; cur.m4 = co * cur.m4 - s * a3;
                cur.m6 = co * a6 + s * cur.m7; cur.m7 = co * cur.m7 - s * a6;
            }
        }
        if (ax == F_AXIS_X) { wx = sg * cur.m0; wy = sg * cur.m3; wz = sg * cur.m6; }
        else if (ax == F_AXIS_Y) { wx = sg * cur.m1; wy = sg * cur.m4; wz = sg * cur.m7; }
        else { wx = sg * cur.m2; wy = sg * cur.m5; wz = sg * cur.m8; }
    }
    if (op.flags & F_PRISMATIC)
    {
        // sn_i carries q for prismatic joints
        cur.x = fma(wx, sn_i, cur.x);
        cur.y = fma(wy, sn_i, cur.y);
        cur.z = fma(wz, sn_i, cur.z);
    }
    W[0] = wx; W[1] = wy; W[2] = wz;
    PJ[0] = cur.x; PJ[1] = cur.y; PJ[2] = cur.z;
}

template <int NJ>
__global__ void __launch_bounds__(CTPB) exo_chain_kernel(const KernelParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    {
        const int n_words = P.program_bytes >> 2;
        const int32_t* src = reinterpret_cast<const int32_t*>(P.program);
        int32_t* dst = reinterpret_cast<int32_t*>(smem);
        for (int i = tid; i < n_words; i += CTPB) dst[i] = src[i];
    }
    __syncthreads();
    const ProgramHeader* H = reinterpret_cast<const ProgramHeader*>(smem);
    const Op* ops = reinterpret_cast<const Op*>(smem + sizeof(ProgramHeader));
    const int32_t* ints = reinterpret_cast<const int32_t*>(ops + H->n_ops);
    const double* dbls = reinterpret_cast<const double*>(smem + P.dbl_offset);
    const int32_t* jpc = ints + H->chain_tab;
    const int N = H->n_ctrl;
    double* qs = reinterpret_cast<double*>(smem + P.scratch_offset) + tid;  // q[i] at qs[i * CTPB]

    for (int64_t e = (int64_t)blockIdx.x * CTPB + tid; e < P.n_evals; e += (int64_t)gridDim.x * CTPB)
    {
        {
            const double* xe = P.x + e * P.x_se;
            for (int i = 0; i < N; ++i) qs[i * CTPB] = xe[i * P.x_sk];
        }
        double* phi_e = P.phi + e * P.phi_se;
        double* jac_e = P.jac ? P.jac + e * P.j_se : nullptr;

        // all joint angles' sin/cos first: NJ independent chains for the FP64 pipe
        double sn[NJ], cs[NJ];
#pragma unroll
        for (int i = 0; i < NJ; ++i)
        {
            const Op& op = ops[jpc[i]];
            const double* c = dbls + op.cst;
            double q = c[CJ_QOFF];
            if (op.qidx >= 0) q = fma(c[CJ_QMULT], qs[op.qidx * CTPB], q);
            if (op.flags & F_PRISMATIC)
            {
                sn[i] = q;
                cs[i] = 0.0;
            }
            else
                sincos(q, &sn[i], &cs[i]);
        }
        // prefix: emits that need no joint (JointLimit, unused-map fills)
        int pc = 0;
        while (ops[pc].kind >= OP_EMIT_FRAME)
        {
            const Op& op = ops[pc];
            const double* c = dbls + op.cst;
            if (op.kind == OP_EMIT_JLIMIT)
                emit_joint_limit(c, N, qs, CTPB, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, jac_e ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            else
                emit_fill(c, op.n_items, op.n_rows, N, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, jac_e ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            ++pc;
        }
        // the chain itself, fully unrolled: frame, joint axes and joint origins stay in registers
        Fr cur = fr_identity();
        double W[NJ][3], PJ[NJ][3];
#pragma unroll
        for (int i = 0; i < NJ; ++i)
        {
            const Op& op = ops[pc + i];
            chain_joint<NJ>(op, dbls + op.cst, cur, sn[i], cs[i], W[i], PJ[i]);
        }
        pc += NJ;
        // suffix: every requested frame hangs off the last joint
        while (ops[pc].kind == OP_EMIT_FRAME)
        {
            const Op& op = ops[pc];
            chain_emit_frame<NJ, NJ>(op, dbls + op.cst, ints + op.tab, N, NJ, cur, W, PJ, phi_e, jac_e, P);
            ++pc;
        }
    }
}

template <int NJ>
cudaError_t chain_plan(size_t smem_bytes, int max_smem, int& ctas)
{
    cudaError_t err = cudaFuncSetAttribute(exo_chain_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, exo_chain_kernel<NJ>, CTPB, smem_bytes);
}
}  // namespace

bool chain_supported(const ProgramHeader& hdr)
{
    return hdr.chain_joints >= 1 && hdr.chain_joints <= 8;
}

cudaError_t plan_chain_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, LaunchConfig& cfg, KernelParams& params)
{
    int dev = 0, sms = 0, max_smem = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (err != cudaSuccess) return err;
    const size_t prog = ((size_t)program_bytes + 15) & ~(size_t)15;
    cfg.threads = CTPB;
    cfg.smem_bytes = prog + (size_t)hdr.n_ctrl * CTPB * sizeof(double);
    params.scratch_offset = (int32_t)prog;
    if (cfg.smem_bytes > (size_t)max_smem) return cudaErrorInvalidConfiguration;
    int ctas = 0;
    switch (hdr.chain_joints)
    {
        case 1: err = chain_plan<1>(cfg.smem_bytes, max_smem, ctas); break;
        case 2: err = chain_plan<2>(cfg.smem_bytes, max_smem, ctas); break;
        case 3: err = chain_plan<3>(cfg.smem_bytes, max_smem, ctas); break;
        case 4: err = chain_plan<4>(cfg.smem_bytes, max_smem, ctas); break;
        case 5: err = chain_plan<5>(cfg.smem_bytes, max_smem, ctas); break;
        case 6: err = chain_plan<6>(cfg.smem_bytes, max_smem, ctas); break;
        case 7: err = chain_plan<7>(cfg.smem_bytes, max_smem, ctas); break;
        case 8: err = chain_plan<8>(cfg.smem_bytes, max_smem, ctas); break;
        default: return cudaErrorInvalidConfiguration;
    }
    if (err != cudaSuccess) return err;
    if (ctas < 1) return cudaErrorInvalidConfiguration;
    const int64_t want = (n_evals + CTPB - 1) / CTPB;
    const int64_t resident = (int64_t)sms * ctas;
    cfg.blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    return cudaSuccess;
}

cudaError_t launch_chain(int chain_joints, const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream)
{
    switch (chain_joints)
    {
        case 1: exo_chain_kernel<1><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 2: exo_chain_kernel<2><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 3: exo_chain_kernel<3><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 4: exo_chain_kernel<4><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 5: exo_chain_kernel<5><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 6: exo_chain_kernel<6><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 7: exo_chain_kernel<7><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        case 8: exo_chain_kernel<8><<<cfg.blocks, CTPB, cfg.smem_bytes, stream>>>(params); break;
        default: return cudaErrorInvalidConfiguration;
    }
    return cudaGetLastError();
}

}  // namespace exo
