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

template <int TPB>
__global__ void __launch_bounds__(TPB) exo_eval_kernel(const KernelParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    // ---- stage the program (topology, axes, offsets, tables) once per CTA ---------------------------------
    {
        const int n_words = P.program_bytes >> 2;
        const int32_t* src = reinterpret_cast<const int32_t*>(P.program);
        int32_t* dst = reinterpret_cast<int32_t*>(smem);
        for (int i = tid; i < n_words; i += TPB) dst[i] = src[i];
    }
    __syncthreads();
    const ProgramHeader* H = reinterpret_cast<const ProgramHeader*>(smem);
    const Op* ops = reinterpret_cast<const Op*>(smem + sizeof(ProgramHeader));
    const int32_t* ints = reinterpret_cast<const int32_t*>(ops + H->n_ops);
    const double* dbls = reinterpret_cast<const double*>(smem + P.dbl_offset);
    double* scr = reinterpret_cast<double*>(smem + P.scratch_offset) + tid;  // word k at scr[k * TPB]
    const int N = H->n_ctrl;
    const int joint_base = 12 * H->n_frame_slots;
    const int acc_base = joint_base + 6 * H->n_joint_slots;
    double* qs = scr + (size_t)H->scratch_doubles * TPB;
    const bool want_j = P.jac != nullptr;

    for (int64_t e = (int64_t)blockIdx.x * TPB + tid; e < P.n_evals; e += (int64_t)gridDim.x * TPB)
    {
        // ---- stage this evaluation's state (N independent loads in flight) --------------------------------
        {
            const double* xe = P.x + e * P.x_se;
            for (int i = 0; i < N; ++i) qs[i * TPB] = xe[i * P.x_sk];
        }
        double* phi_e = P.phi + e * P.phi_se;
        double* jac_e = want_j ? P.jac + e * P.j_se : nullptr;
        Fr cur = fr_identity();

        for (int pc = 0;; ++pc)
        {
            const Op op = ops[pc];
            if (op.kind == OP_END) break;
            const double* c = dbls + op.cst;
            if (op.kind == OP_JOINT || op.kind == OP_TABLE)
            {
                if (op.src == SRC_ROOT)
                    cur = fr_identity();
                else if (op.src >= 0)
                    cur = fr_load_scratch<TPB>(scr + 12 * op.src * TPB);
                const int pre = op.flags & F_PRE_MASK;
                if (pre == F_PRE_TRANS)
                    fr_translate(cur, c[9], c[10], c[11]);
                else if (pre == F_PRE_FULL)
                    cur = fr_mul(cur, fr_load_const(c));
                if (op.kind == OP_TABLE)
                {
                    const int t = (int)(e % P.T);
                    const double* g = P.generated + ((size_t)t * H->n_generated + op.qidx) * 12;
                    Fr gf;
                    gf.m0 = g[0]; gf.m1 = g[1]; gf.m2 = g[2]; gf.m3 = g[3]; gf.m4 = g[4]; gf.m5 = g[5];
                    gf.m6 = g[6]; gf.m7 = g[7]; gf.m8 = g[8]; gf.x = g[9]; gf.y = g[10]; gf.z = g[11];
                    cur = fr_mul(cur, gf);
                }
                else
                {
                    double q = c[CJ_QOFF];
                    if (op.qidx >= 0) q = fma(c[CJ_QMULT], qs[op.qidx * TPB], q);
                    const int ax = op.flags & F_AXIS_MASK;
                    double wx, wy, wz;  // axis in world coordinates
                    if (op.flags & F_PRISMATIC)
                    {
                        if (ax == F_AXIS_GENERAL)
                        {
                            const double ux = c[CJ_AXIS], uy = c[CJ_AXIS + 1], uz = c[CJ_AXIS + 2];
                            wx = cur.m0 * ux + cur.m1 * uy + cur.m2 * uz;
                            wy = cur.m3 * ux + cur.m4 * uy + cur.m5 * uz;
                            wz = cur.m6 * ux + cur.m7 * uy + cur.m8 * uz;
                        }
                        else
                        {
                            const double sg = (op.flags & F_AXIS_NEG) ? -1.0 : 1.0;
                            if (ax == F_AXIS_X) { wx = sg * cur.m0; wy = sg * cur.m3; wz = sg * cur.m6; }
                            else if (ax == F_AXIS_Y) { wx = sg * cur.m1; wy = sg * cur.m4; wz = sg * cur.m7; }
                            else { wx = sg * cur.m2; wy = sg * cur.m5; wz = sg * cur.m8; }
                        }
                        cur.x = fma(wx, q, cur.x);
                        cur.y = fma(wy, q, cur.y);
                        cur.z = fma(wz, q, cur.z);
                    }
                    else
                    {
                        double s, cs;
                        sincos(q, &s, &cs);
                        if (ax == F_AXIS_GENERAL)
                        {
                            const double ux = c[CJ_AXIS], uy = c[CJ_AXIS + 1], uz = c[CJ_AXIS + 2];
                            const double vt = 1.0 - cs;
                            Fr r;
                            r.m0 = cs + vt * ux * ux;
                            r.m1 = -uz * s + vt * ux * uy;
                            r.m2 = uy * s + vt * ux * uz;
                            r.m3 = uz * s + vt * ux * uy;
                            r.m4 = cs + vt * uy * uy;
                            r.m5 = -ux * s + vt * uy * uz;
                            r.m6 = -uy * s + vt * ux * uz;
                            r.m7 = ux * s + vt * uy * uz;
                            r.m8 = cs + vt * uz * uz;
                            r.x = r.y = r.z = 0.0;
                            const double px = cur.x, py = cur.y, pz = cur.z;
                            cur = fr_mul(cur, r);
                            cur.x = px; cur.y = py; cur.z = pz;
                            wx = cur.m0 * ux + cur.m1 * uy + cur.m2 * uz;
                            wy = cur.m3 * ux + cur.m4 * uy + cur.m5 * uz;
                            wz = cur.m6 * ux + cur.m7 * uy + cur.m8 * uz;
                        }
                        else
                        {
                            const bool neg = (op.flags & F_AXIS_NEG) != 0;
                            if (neg) s = -s;
                            const double sg = neg ? -1.0 : 1.0;
                            if (ax == F_AXIS_X)
                            {
                                // M * Rx: col1' = c col1 + s col2, col2' = -s col1 + c col2
                                const double a1 = cur.m1, a4 = cur.m4, a7 = cur.m7;
                                cur.m1 = cs * a1 + s * cur.m2; cur.m2 = cs * cur.m2 - s * a1;
                                cur.m4 = cs * a4 + s * cur.m5; cur.m5 = cs * cur.m5 - s * a4;
                                cur.m7 = cs * a7 + s * cur.m8; cur.m8 = cs * cur.m8 - s * a7;
                                wx = sg * cur.m0; wy = sg * cur.m3; wz = sg * cur.m6;
                            }
                            else if (ax == F_AXIS_Y)
                            {
                                // M * Ry: col0' = c col0 - s col2, col2' = s col0 + c col2
                                const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                                cur.m0 = cs * a0 - s * cur.m2; cur.m2 = s * a0 + cs * cur.m2;
                                cur.m3 = cs * a3 - s * cur.m5; cur.m5 = s * a3 + cs * cur.m5;
                                cur.m6 = cs * a6 - s * cur.m8; cur.m8 = s * a6 + cs * cur.m8;
                                wx = sg * cur.m1; wy = sg * cur.m4; wz = sg * cur.m7;
                            }
                            else
                            {
                                // M * Rz: col0' = c col0 + s col1, col1' = -s col0 + c col1
                                const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                                cur.m0 = cs * a0 + s * cur.m1; cur.m1 = cs * cur.m1 - s * a0;
                                cur.m3 = cs * a3 + s * cur.m4; cur.m4 = cs * cur.m4 - s * a3;
                                cur.m6 = cs * a6 + s * cur.m7; cur.m7 = cs * cur.m7 - s * a6;
                                wx = sg * cur.m2; wy = sg * cur.m5; wz = sg * cur.m8;
                            }
                        }
                    }
                    if (op.jslot >= 0)
                    {
                        double* js = scr + (joint_base + 6 * op.jslot) * TPB;
                        js[0] = wx; js[TPB] = wy; js[2 * TPB] = wz;
                        js[3 * TPB] = cur.x; js[4 * TPB] = cur.y; js[5 * TPB] = cur.z;
                    }
                }
                if (op.dst >= 0) fr_store_scratch<TPB>(scr + 12 * op.dst * TPB, cur);
            }
            else if (op.kind == OP_EMIT_FRAME)
            {
                // temp_A = frame_A * offA ; temp_B = frame_B * offB ; Phi = temp_B^-1 * temp_A  (kinematic_tree.cpp:829-835)
                Fr A;
                if (op.src == SRC_CUR) A = cur;
                else if (op.src == SRC_ROOT) A = fr_identity();
                else A = fr_load_scratch<TPB>(scr + 12 * op.src * TPB);
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
                        else B = fr_load_scratch<TPB>(scr + 12 * op.src_b * TPB);
                        B = fr_mul(B, fr_load_const(c + CE_B_M));
                        Binv = fr_inverse(B);
                    }
                    Phi = fr_mul(Binv, A);
                }
                write_phi(op.map_kind, op.rot_type, Phi, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk);
                if (want_j)
                {
                    // ComputeJ (kinematic_tree.cpp:919-945): column c = temp_B.M^-1 * [ w x (temp_A.p - p_joint) ; w ]
                    const int32_t* tab = ints + op.tab;
                    double* jr = jac_e + (int64_t)op.jrow_off * P.j_sr;
                    const int64_t sr = P.j_sr;
                    const bool three = op.n_rows == 3;
                    for (int col = 0; col < N; ++col)
                    {
                        const int ent = tab[col];
                        double v0 = 0.0, v1 = 0.0, v2 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
                        if (ent >= 0)
                        {
                            const double* js = scr + (joint_base + 6 * (ent >> 2)) * TPB;
                            const double ax = js[0], ay = js[TPB], az = js[2 * TPB];
                            if (ent & 2)
                            {
                                v0 = ax; v1 = ay; v2 = az;
                            }
                            else
                            {
                                const double dx = A.x - js[3 * TPB], dy = A.y - js[4 * TPB], dz = A.z - js[5 * TPB];
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
                        double* jc = jr + (int64_t)col * P.j_sc;
                        jc[0] = v0; jc[sr] = v1; jc[2 * sr] = v2;
                        if (!three)
                        {
                            jc[3 * sr] = w0; jc[4 * sr] = w1; jc[5 * sr] = w2;
                        }
                    }
                }
            }
            else if (op.kind == OP_EMIT_JLIMIT)
            {
                emit_joint_limit(c, N, qs, TPB, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, want_j ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            }
            else if (op.kind == OP_EMIT_FILL)
            {
                emit_fill(c, op.n_items, op.n_rows, N, phi_e + (int64_t)op.phi_off * P.phi_sk, P.phi_sk, want_j ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr, P.j_sr, P.j_sc);
            }
            else if (op.kind == OP_EMIT_SPHERES)
            {
                // SphereCollision::Update (sphere_collision.cpp:123-162).  The reference accumulates, per sphere pair
                // and per column, -(J_i - J_j)^T d / |d|; here the pair loop accumulates unit "forces" per carrying
                // link (F, M = p x F) and the Jacobian row is recovered per joint as  -w . (sum M - p_joint x sum F)
                // (same sum, regrouped: O(pairs + joints) instead of O(pairs * joints)).
                const int32_t* tab = ints + op.tab;
                const int G = tab[0], n_acc = tab[1];
                const int32_t* gs = tab + 2;
                const int32_t* sph = gs + G + 1;
                const int32_t* cols = sph + SPH_INTS * op.n_items;
                const double k5 = c[0];
                const double* sc = c + 1;  // per sphere: ox oy oz radius
                double* acc = scr + (size_t)acc_base * TPB;
                int out_idx = 0;
                for (int ga = 0; ga < G; ++ga)
                {
                    for (int gb = ga + 1; gb < G; ++gb, ++out_idx)
                    {
                        for (int a = 0; a < 6 * n_acc; ++a) acc[a * TPB] = 0.0;
                        double phi_sum = 0.0;
                        for (int i = gs[ga]; i < gs[ga + 1]; ++i)
                        {
                            const int src_i = sph[SPH_INTS * i], acc_i = sph[SPH_INTS * i + 1];
                            double pix = sc[4 * i], piy = sc[4 * i + 1], piz = sc[4 * i + 2];
                            const double ri = sc[4 * i + 3];
                            if (src_i != SRC_ROOT)
                            {
                                const Fr f = (src_i == SRC_CUR) ? cur : fr_load_scratch<TPB>(scr + 12 * src_i * TPB);
                                const double tx = f.m0 * pix + f.m1 * piy + f.m2 * piz + f.x;
                                const double ty = f.m3 * pix + f.m4 * piy + f.m5 * piz + f.y;
                                const double tz = f.m6 * pix + f.m7 * piy + f.m8 * piz + f.z;
                                pix = tx; piy = ty; piz = tz;
                            }
                            double gx = 0.0, gy = 0.0, gz = 0.0;
                            for (int j = gs[gb]; j < gs[gb + 1]; ++j)
                            {
                                const int src_j = sph[SPH_INTS * j], acc_j = sph[SPH_INTS * j + 1];
                                double pjx = sc[4 * j], pjy = sc[4 * j + 1], pjz = sc[4 * j + 2];
                                if (src_j != SRC_ROOT)
                                {
                                    const Fr f = (src_j == SRC_CUR) ? cur : fr_load_scratch<TPB>(scr + 12 * src_j * TPB);
                                    const double tx = f.m0 * pjx + f.m1 * pjy + f.m2 * pjz + f.x;
                                    const double ty = f.m3 * pjx + f.m4 * pjy + f.m5 * pjz + f.y;
                                    const double tz = f.m6 * pjx + f.m7 * pjy + f.m8 * pjz + f.z;
                                    pjx = tx; pjy = ty; pjz = tz;
                                }
                                const double dx = pix - pjx, dy = piy - pjy, dz = piz - pjz;
                                const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
                                phi_sum += 1.0 / (1.0 + exp(k5 * (nrm - ri - sc[4 * j + 3])));
                                if (want_j)
                                {
                                    const double ux = dx / nrm, uy = dy / nrm, uz = dz / nrm;
                                    gx += ux; gy += uy; gz += uz;
                                    if (acc_j >= 0)
                                    {
                                        double* aj = acc + 6 * acc_j * TPB;
                                        aj[0] -= ux; aj[TPB] -= uy; aj[2 * TPB] -= uz;
                                        aj[3 * TPB] -= pjy * uz - pjz * uy;
                                        aj[4 * TPB] -= pjz * ux - pjx * uz;
                                        aj[5 * TPB] -= pjx * uy - pjy * ux;
                                    }
                                }
                            }
                            if (want_j && acc_i >= 0)
                            {
                                double* ai = acc + 6 * acc_i * TPB;
                                ai[0] += gx; ai[TPB] += gy; ai[2 * TPB] += gz;
                                ai[3 * TPB] += piy * gz - piz * gy;
                                ai[4 * TPB] += piz * gx - pix * gz;
                                ai[5 * TPB] += pix * gy - piy * gx;
                            }
                        }
                        phi_e[(int64_t)(op.phi_off + out_idx) * P.phi_sk] = phi_sum;
                        if (want_j)
                        {
                            double* row = jac_e + (int64_t)(op.jrow_off + out_idx) * P.j_sr;
                            for (int col = 0; col < N; ++col)
                            {
                                const int ent = cols[3 * col];
                                double val = 0.0;
                                if (ent >= 0)
                                {
                                    const uint32_t mlo = (uint32_t)cols[3 * col + 1], mhi = (uint32_t)cols[3 * col + 2];
                                    double fx = 0, fy = 0, fz = 0, mx = 0, my = 0, mz = 0;
                                    for (int a = 0; a < n_acc; ++a)
                                    {
                                        const bool in = a < 32 ? ((mlo >> a) & 1u) : ((mhi >> (a - 32)) & 1u);
                                        if (!in) continue;
                                        const double* aa = acc + 6 * a * TPB;
                                        fx += aa[0]; fy += aa[TPB]; fz += aa[2 * TPB];
                                        mx += aa[3 * TPB]; my += aa[4 * TPB]; mz += aa[5 * TPB];
                                    }
                                    const double* js = scr + (joint_base + 6 * (ent >> 2)) * TPB;
                                    const double ax = js[0], ay = js[TPB], az = js[2 * TPB];
                                    if (ent & 2)
                                        val = -(ax * fx + ay * fy + az * fz);
                                    else
                                    {
                                        const double px = js[3 * TPB], py = js[4 * TPB], pz = js[5 * TPB];
                                        const double tx = mx - (py * fz - pz * fy);
                                        const double ty = my - (pz * fx - px * fz);
                                        const double tz = mz - (px * fy - py * fx);
                                        val = -(ax * tx + ay * ty + az * tz);
                                    }
                                }
                                row[(int64_t)col * P.j_sc] = val;
                            }
                        }
                    }
                }
            }
        }
    }
}
}  // namespace

template <int TPB>
cudaError_t try_config(size_t prog, size_t per_thread, int max_smem, int& ctas_per_sm, size_t& smem_bytes)
{
    smem_bytes = prog + per_thread * TPB;
    ctas_per_sm = 0;
    if (smem_bytes > (size_t)max_smem) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(exo_eval_kernel<TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, exo_eval_kernel<TPB>, TPB, smem_bytes);
}

cudaError_t plan_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, LaunchConfig& cfg, KernelParams& params)
{
    int dev = 0, sms = 0, max_smem = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (err != cudaSuccess) return err;
    const size_t prog = ((size_t)program_bytes + 15) & ~(size_t)15;
    const size_t per_thread = (size_t)(hdr.scratch_doubles + hdr.n_ctrl) * sizeof(double);
    params.scratch_offset = (int32_t)prog;
    // pick the CTA size that keeps the most threads resident per SM (per-thread scratch lives in shared memory)
    int best_threads = 0, best_tpb = 0, best_ctas = 0;
    size_t best_smem = 0;
    const int sizes[3] = {128, 64, 32};
    for (int k = 0; k < 3; ++k)
    {
        int ctas = 0;
        size_t smem_bytes = 0;
        if (sizes[k] == 128) err = try_config<128>(prog, per_thread, max_smem, ctas, smem_bytes);
        else if (sizes[k] == 64) err = try_config<64>(prog, per_thread, max_smem, ctas, smem_bytes);
        else err = try_config<32>(prog, per_thread, max_smem, ctas, smem_bytes);
        if (err != cudaSuccess) return err;
        if (ctas * sizes[k] > best_threads)
        {
            best_threads = ctas * sizes[k];
            best_tpb = sizes[k];
            best_ctas = ctas;
            best_smem = smem_bytes;
        }
    }
    if (best_threads == 0) return cudaErrorInvalidConfiguration;
    cfg.threads = best_tpb;
    cfg.smem_bytes = best_smem;
    const int64_t want = (n_evals + best_tpb - 1) / best_tpb;
    const int64_t resident = (int64_t)sms * best_ctas;
    cfg.blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    return cudaSuccess;
}

cudaError_t launch_eval(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream)
{
    if (cfg.threads == 128) exo_eval_kernel<128><<<cfg.blocks, 128, cfg.smem_bytes, stream>>>(params);
    else if (cfg.threads == 64) exo_eval_kernel<64><<<cfg.blocks, 64, cfg.smem_bytes, stream>>>(params);
    else exo_eval_kernel<32><<<cfg.blocks, 32, cfg.smem_bytes, stream>>>(params);
    return cudaGetLastError();
}

}  // namespace exo
