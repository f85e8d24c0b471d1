// capi.cu -- implementation of the C-ABI declared in include/exotica_b200.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/exotica_b200.h"
#include "compiler.h"
#include "kernels.cuh"

using namespace exo;

namespace
{
thread_local std::string g_error;

int fail(int code, const std::string& msg)
{
    g_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    return EXO_ERR_CUDA;
}

#define CUDA_TRY(expr)                                        \
    do                                                        \
    {                                                         \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);   \
    } while (0)
}  // namespace

struct exo_tree
{
    TreeModel model;
    int64_t version = 0;  // bumped when model state / limits change, so problems recompile
};

#ifndef EXO_PIPE_NBUF
#define EXO_PIPE_NBUF 8  // chunks in flight (measured on cfg2a, 64 MiB chunks: 3 -> 8.35 ms, 4 -> 8.03, 6 -> 8.05, 8 -> 7.96 ms per 461 MB step); in the host-buffer pipelines (one stream and one buffer set each)
#endif

struct exo_problem
{
    exo_tree* tree = nullptr;
    std::vector<MapSpec> maps;
    bool finalized = false;
    int64_t compiled_version = -1;
    Program prog;
    unsigned char* d_image = nullptr;
    size_t d_image_bytes = 0;
    // per-timestep generated frames
    int T = 1, n_generated = 0;
    std::vector<double> h_generated;
    double* d_generated = nullptr;
    // raw-frames shadow problem behind exo_problem_hessian_batch (all frames of all maps, full 6 x N Jacobians)
    exo_problem* shadow = nullptr;
    std::vector<int> shadow_frame_start;  // per map
    double *d_shadow_phi = nullptr, *d_shadow_jac = nullptr;
    int64_t shadow_capacity = 0;
    // batched IK workspace (exo_problem_ik_solve_batch)
    double* d_ik = nullptr;     // doubles: phi, jac, q_trial, qd, err_prev, stop, lambda
    int32_t* d_ik_i = nullptr;  // ints: active, ls_done, n_active, n_list, list
    int64_t ik_capacity = 0;
    std::vector<double> cost_rho_t0;  // rho of every map at t = 0 (IK uses end-pose problems)
    std::vector<double> cost_rho;     // host copy of rho[T][n_maps] (eligibility of the fused cost-gradient kernel)
    double* d_cg = nullptr;           // scratch of exo_problem_cost_gradient_batch: per-evaluation costs (+ Phi, J when not fused)
    int64_t cg_bytes = 0;
    std::vector<double> cost_W;
    // auxiliary outputs and the task cost (rows e / f1)
    double* d_min_dist = nullptr;  // caller-owned, [n_evals] (exo_problem_set_min_distance_output)
    bool cost_set = false;
    CostParams cost;               // constants of the cost kernel; device copies below
    double *d_rho = nullptr, *d_goal = nullptr, *d_goal_rot = nullptr, *d_W = nullptr;
    // timing / accounting
    int64_t launches = 0;
    int compile_mode = COMPILE_AUTO;  // exo_problem_set_kernel choices 4 / 5 force the interpreter's one- / two-pass program
    int kernel_choice = 0;        // 0 auto, 1 interpreter only, 2..3 chain kernel variants (CHAIN_REGS, CHAIN_TWO_PASS)
    bool chain_ok = false;        // the compiled program lowered to ChainParams
    ChainParams chain;            // problem constants of the chain kernel (buffers / strides filled per launch)
    // TMA output path of the chain kernel: staging plan + tensor maps of the last buffers bound (re-encoded when they change)
    ChainTmaParams chain_tma;
    struct TmaKey
    {
        const void *phi = nullptr, *jac = nullptr;
        int64_t n = 0, phi_sk = 0, j_sc = 0, j_sr = 0;
        int tile = -1, version = -1;
        bool ok = false;
    } chain_tma_key;
    // per-robot generated kernel for tree programs (tree_codegen.cpp / tree_jit.cu)
    int tree_small_calls = 0;     // launches with 256 <= E < EXO_TREE_MIN_EVALS so far: from the 16th on they take the generated kernel too
    int tree_version = -2;        // compiled_version the plan below was generated for
    bool tree_ok = false;         // source generated (the program is inside the generator's scope)
    void* tree_kernel = nullptr;  // loaded module entry (nullptr until the first eligible launch on a device)
    bool tree_jit_failed = false;
    std::string tree_why;         // why the generated kernel does not apply
    TreeGenPlan tree_plan;
    TreeGenParams tree_params;
    TmaKey tree_key;
    // generated front half of split SphereCollision programs (tree_park_codegen)
    int park_version = -2;
    bool park_ok = false, park_jit_failed = false;
    void* park_kernel = nullptr;
    TreeParkPlan park_plan;
    // generated fused cost + gradient kernel for tree programs (tree_costgrad_codegen)
    int tcg_version = -2;         // tree version the source below was generated for
    int64_t cost_epoch = 0, tcg_cost_epoch = -1;  // exo_problem_set_cost calls / the one the cost structure below was read at
    bool tcg_ok = false;
    void* tcg_kernel = nullptr;
    bool tcg_jit_failed = false;
    std::string tcg_why;
    TreeCostSpec tcg_spec;
    TreeCostPlan tcg_plan;
    // generated update + cost kernel (tree_codegen with a cost epilogue) behind exo_problem_update_cost_batch
    int uc_version = -2;
    int64_t uc_cost_epoch = -1;
    bool uc_ok = false, uc_jit_failed = false;
    void* uc_kernel = nullptr;
    std::string uc_why;
    TreeCostSpec uc_spec;
    TreeGenPlan uc_plan;
    TreeGenParams uc_params;
    TmaKey uc_key;
    double* d_uc_cost = nullptr;  // per-evaluation costs when the caller does not ask for them
    int64_t uc_cost_bytes = 0;
    bool plan_valid = false;      // interpreter launch plan (depends on the program and on split_spheres only)
    bool plan_split = false;
    bool fused_spheres = false;   // exo_problem_set_kernel(6): SphereCollision inside the interpreter (no second launch)
    struct SphDump                // slots parked between the interpreter and exo_spheres_kernel: one buffer per stream in use
    {
        cudaStream_t stream;
        double* d;
        int64_t bytes;
    };
    std::vector<SphDump> sph_dumps;
    LaunchConfig plan_cfg;
    int32_t plan_scratch_offset = 0;
    const char* last_kernel = "none";
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // host pipeline
    static const int NBUF = EXO_PIPE_NBUF;
    cudaStream_t streams[NBUF] = {};
    double* d_x[NBUF] = {};
    double* d_phi[NBUF] = {};
    double* d_jac[NBUF] = {};
    double* d_cost[NBUF] = {};
    double* d_grad[NBUF] = {};
    double* d_cgs[NBUF] = {};       // per-stream scratch of the cost-gradient chunks (cost_gradient_impl)
    int64_t cgs_bytes[NBUF] = {};
    int64_t pipe_chunk = 0;
    bool pipe_has_jac = false;
};

extern "C" {

// used by host/host_capi.cpp to report exotica::Exception messages through the same channel
void exo_internal_set_error(const char* msg)
{
    g_error = msg ? msg : "";
}

const char* exo_last_error(void)
{
    return g_error.c_str();
}

const char* exo_version(void)
{
    return "exotica_b200 0.1 (sm_100a)";
}

int exo_device_count(int* count)
{
    if (!count) return fail(EXO_ERR_INVALID, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess)
    {
        *count = 0;
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    return EXO_OK;
}

int exo_set_device(int device)
{
    CUDA_TRY(cudaSetDevice(device));
    return EXO_OK;
}

int exo_device_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return fail(EXO_ERR_INVALID, "ptr is NULL");
    CUDA_TRY(cudaMalloc(ptr, bytes ? bytes : 8));
    return EXO_OK;
}

int exo_device_free(void* ptr)
{
    CUDA_TRY(cudaFree(ptr));
    return EXO_OK;
}

int exo_host_alloc_pinned(void** ptr, size_t bytes)
{
    if (!ptr) return fail(EXO_ERR_INVALID, "ptr is NULL");
    CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 8));
    return EXO_OK;
}

int exo_host_free_pinned(void* ptr)
{
    CUDA_TRY(cudaFreeHost(ptr));
    return EXO_OK;
}

int exo_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return EXO_OK;
}

int exo_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return EXO_OK;
}

int exo_memset_device(void* dst, int value, size_t bytes, void* stream)
{
    CUDA_TRY(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return EXO_OK;
}

int exo_stream_create(void** stream)
{
    if (!stream) return fail(EXO_ERR_INVALID, "stream is NULL");
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return EXO_OK;
}

int exo_stream_destroy(void* stream)
{
    CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream));
    return EXO_OK;
}

int exo_stream_synchronize(void* stream)
{
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return EXO_OK;
}

int exo_device_synchronize(void)
{
    CUDA_TRY(cudaDeviceSynchronize());
    return EXO_OK;
}

// ---- tree ---------------------------------------------------------------------------------------------
int exo_tree_create(const exo_tree_desc* d, exo_tree** out)
{
    if (!d || !out) return fail(EXO_ERR_INVALID, "NULL argument");
    if (d->n_elements < 1 || !d->id || !d->parent || !d->joint_type || !d->joint_origin || !d->joint_axis || !d->f_tip || !d->control_id)
        return fail(EXO_ERR_INVALID, "incomplete tree description");
    if (d->n_controlled < 1) return fail(EXO_ERR_INVALID, "No update joints specified!");
    exo_tree* t = new exo_tree();
    TreeModel& m = t->model;
    const int n = d->n_elements;
    m.n = n;
    m.n_ctrl = d->n_controlled;
    m.id.assign(d->id, d->id + n);
    m.parent.assign(d->parent, d->parent + n);
    m.joint_type.assign(d->joint_type, d->joint_type + n);
    m.control_id.assign(d->control_id, d->control_id + n);
    m.origin.assign(d->joint_origin, d->joint_origin + 3 * (size_t)n);
    m.axis.assign(d->joint_axis, d->joint_axis + 3 * (size_t)n);
    m.tip.assign(d->f_tip, d->f_tip + 12 * (size_t)n);
    m.mimic_elem.assign(n, -1);
    m.mimic_mult.assign(n, 1.0);
    m.mimic_off.assign(n, 0.0);
    m.generated.assign(n, -1);
    if (d->mimic_element) m.mimic_elem.assign(d->mimic_element, d->mimic_element + n);
    if (d->mimic_multiplier) m.mimic_mult.assign(d->mimic_multiplier, d->mimic_multiplier + n);
    if (d->mimic_offset) m.mimic_off.assign(d->mimic_offset, d->mimic_offset + n);
    if (d->generated_slot) m.generated.assign(d->generated_slot, d->generated_slot + n);
    m.n_generated = 0;
    for (int e = 0; e < n; ++e) m.n_generated = std::max(m.n_generated, m.generated[e] + 1);
    m.tree_state.assign(n, 0.0);
    m.lower.assign(m.n_ctrl, -1e20);
    m.upper.assign(m.n_ctrl, 1e20);
    if (d->lower) m.lower.assign(d->lower, d->lower + m.n_ctrl);
    if (d->upper) m.upper.assign(d->upper, d->upper + m.n_ctrl);
    std::vector<char> seen(m.n_ctrl, 0);
    for (int e = 0; e < n; ++e)
    {
        if (m.joint_type[e] < 0 || m.joint_type[e] > EXO_JOINT_NONE)
        {
            delete t;
            return fail(EXO_ERR_INVALID, "unknown joint type");
        }
        if (m.parent[e] >= e || (e > 0 && m.parent[e] < 0) || (e == 0 && m.parent[e] != -1))
        {
            delete t;
            return fail(EXO_ERR_INVALID, "tree elements must be ordered parents-first with the root at index 0");
        }
        if (m.control_id[e] >= m.n_ctrl)
        {
            delete t;
            return fail(EXO_ERR_INVALID, "control_id out of range");
        }
        if (m.control_id[e] >= 0) seen[m.control_id[e]] = 1;
    }
    for (int c = 0; c < m.n_ctrl; ++c)
        if (!seen[c])
        {
            delete t;
            return fail(EXO_ERR_INVALID, "controlled joint " + std::to_string(c) + " has no element");
        }
    *out = t;
    return EXO_OK;
}

int exo_tree_destroy(exo_tree* tree)
{
    delete tree;
    return EXO_OK;
}

int exo_tree_set_model_state(exo_tree* tree, const double* s)
{
    if (!tree || !s) return fail(EXO_ERR_INVALID, "NULL argument");
    tree->model.tree_state.assign(s, s + tree->model.n);
    tree->version++;
    return EXO_OK;
}

int exo_tree_set_joint_limits(exo_tree* tree, const double* lower, const double* upper)
{
    if (!tree || !lower || !upper) return fail(EXO_ERR_INVALID, "NULL argument");
    tree->model.lower.assign(lower, lower + tree->model.n_ctrl);
    tree->model.upper.assign(upper, upper + tree->model.n_ctrl);
    tree->version++;
    return EXO_OK;
}

int exo_tree_num_controlled(const exo_tree* tree)
{
    return tree ? tree->model.n_ctrl : -1;
}

// ---- problem ------------------------------------------------------------------------------------------------
int exo_problem_create(exo_tree* tree, exo_problem** out)
{
    if (!tree || !out) return fail(EXO_ERR_INVALID, "NULL argument");
    exo_problem* p = new exo_problem();
    p->tree = tree;
    *out = p;
    return EXO_OK;
}

static void free_pipeline(exo_problem* p)
{
    for (int i = 0; i < exo_problem::NBUF; ++i)
    {
        if (p->d_x[i]) cudaFree(p->d_x[i]);
        if (p->d_phi[i]) cudaFree(p->d_phi[i]);
        if (p->d_jac[i]) cudaFree(p->d_jac[i]);
        if (p->d_cost[i]) cudaFree(p->d_cost[i]);
        if (p->d_grad[i]) cudaFree(p->d_grad[i]);
        if (p->d_cgs[i]) cudaFree(p->d_cgs[i]);
        p->d_x[i] = p->d_phi[i] = p->d_jac[i] = p->d_cost[i] = p->d_grad[i] = p->d_cgs[i] = nullptr;
        p->cgs_bytes[i] = 0;
    }
    p->pipe_chunk = 0;
}

int exo_problem_destroy(exo_problem* p)
{
    if (!p) return EXO_OK;
    free_pipeline(p);
    for (int i = 0; i < exo_problem::NBUF; ++i)
        if (p->streams[i]) cudaStreamDestroy(p->streams[i]);
    if (p->d_image) cudaFree(p->d_image);
    if (p->d_generated) cudaFree(p->d_generated);
    if (p->d_ik) cudaFree(p->d_ik);
    if (p->d_ik_i) cudaFree(p->d_ik_i);
    for (auto& sd : p->sph_dumps)
        if (sd.d) cudaFree(sd.d);
    if (p->d_cg) cudaFree(p->d_cg);
    if (p->d_uc_cost) cudaFree(p->d_uc_cost);
    if (p->shadow) exo_problem_destroy(p->shadow);
    if (p->d_shadow_phi) cudaFree(p->d_shadow_phi);
    if (p->d_shadow_jac) cudaFree(p->d_shadow_jac);
    if (p->d_rho) cudaFree(p->d_rho);
    if (p->d_goal) cudaFree(p->d_goal);
    if (p->d_goal_rot) cudaFree(p->d_goal_rot);
    if (p->d_W) cudaFree(p->d_W);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
    return EXO_OK;
}

int exo_problem_add_task_map(exo_problem* p, const exo_task_map_desc* d)
{
    if (!p || !d) return fail(EXO_ERR_INVALID, "NULL argument");
    if (p->finalized) return fail(EXO_ERR_STATE, "problem already finalized");
    if (d->kind < EXO_MAP_EFF_FRAME || d->kind > EXO_MAP_CENTER_OF_MASS) return fail(EXO_ERR_INVALID, "unknown task map kind");
    if (d->n_frames < 0) return fail(EXO_ERR_INVALID, "negative frame count");
    MapSpec m;
    m.kind = d->kind;
    m.n_frames = d->n_frames;
    m.rot_type = d->rotation_type;
    m.safe_percentage = d->safe_percentage;
    m.precision = d->precision;
    m.is_used = d->is_used != 0;
    const int n = p->tree->model.n;
    if (d->n_frames > 0)
    {
        if (!d->elem_a) return fail(EXO_ERR_INVALID, "elem_a is NULL");
        m.elem_a.assign(d->elem_a, d->elem_a + d->n_frames);
        m.elem_b.assign(d->n_frames, 0);
        if (d->elem_b) m.elem_b.assign(d->elem_b, d->elem_b + d->n_frames);
        static const double ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
        m.off_a.resize(12 * (size_t)d->n_frames);
        m.off_b.resize(12 * (size_t)d->n_frames);
        for (int f = 0; f < d->n_frames; ++f)
        {
            memcpy(&m.off_a[12 * f], d->off_a ? d->off_a + 12 * f : ident, sizeof(ident));
            memcpy(&m.off_b[12 * f], d->off_b ? d->off_b + 12 * f : ident, sizeof(ident));
            if (m.elem_a[f] < 0 || m.elem_a[f] >= n) return fail(EXO_ERR_INVALID, "No frame_A link exists with index " + std::to_string(m.elem_a[f]));
            if (m.elem_b[f] < 0 || m.elem_b[f] >= n) return fail(EXO_ERR_INVALID, "No frame_B link exists with index " + std::to_string(m.elem_b[f]));
        }
    }
    if ((m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION) && rotation_length(m.rot_type) < 0)
        return fail(EXO_ERR_INVALID, "Unknown rotation representation type!");
    if (m.kind == EXO_MAP_JOINT_POSE)
    {
        // JointPose::Initialize (joint_pose.cpp:66-97): JointMap defaults to all controlled joints, JointRef to zeros
        const int N = p->tree->model.n_ctrl;
        if (d->n_joint_map < 0) return fail(EXO_ERR_INVALID, "negative JointMap size");
        if (d->n_joint_map > 0)
        {
            if (!d->joint_map) return fail(EXO_ERR_INVALID, "joint_map is NULL");
            m.joint_map.assign(d->joint_map, d->joint_map + d->n_joint_map);
            for (int v : m.joint_map)
                if (v < 0 || v >= N) return fail(EXO_ERR_INVALID, "JointMap index (" + std::to_string(v) + ") is greater than number of controlled joints (" + std::to_string(N) + ")!");
        }
        else
            for (int i = 0; i < N; ++i) m.joint_map.push_back(i);
        m.joint_ref.assign(m.joint_map.size(), 0.0);
        if (d->joint_ref) m.joint_ref.assign(d->joint_ref, d->joint_ref + m.joint_map.size());
    }
    if (m.kind == EXO_MAP_EFF_AXIS_ALIGNMENT)
    {
        // EffAxisAlignment::Initialize (eff_axis_alignment.cpp:39-58): Axis and Direction default to (1, 0, 0) and are normalised
        m.axis.assign(3 * (size_t)d->n_frames, 0.0);
        m.direction.assign(3 * (size_t)d->n_frames, 0.0);
        for (int f = 0; f < d->n_frames; ++f)
        {
            const double unit_x[3] = {1.0, 0.0, 0.0};
            const double* a = d->axis ? d->axis + 3 * f : unit_x;
            const double* b = d->direction ? d->direction + 3 * f : unit_x;
            const double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
            for (int k = 0; k < 3; ++k)
            {
                m.axis[3 * f + k] = a[k] / na;  // Eigen's normalized(): division by the norm (NaN for a zero vector, like the reference)
                m.direction[3 * f + k] = b[k] / nb;
            }
        }
    }
    if (m.kind == EXO_MAP_CENTER_OF_MASS)
    {
        if (d->n_frames > 0 && !d->mass) return fail(EXO_ERR_INVALID, "CenterOfMass needs mass[]");
        m.mass.assign(d->mass, d->mass + d->n_frames);
        m.enable_z = d->enable_z != 0;
    }
    if (m.kind == EXO_MAP_SPHERE_COLLISION)
    {
        if (!d->radius || !d->group) return fail(EXO_ERR_INVALID, "SphereCollision needs radius[] and group[]");
        if (!(d->precision != 0.0)) return fail(EXO_ERR_INVALID, "SphereCollision Precision must be non-zero");
        m.radius.assign(d->radius, d->radius + d->n_frames);
        m.group.assign(d->group, d->group + d->n_frames);
        for (int g : m.group)
            if (g < 0) return fail(EXO_ERR_INVALID, "negative sphere group");
    }
    p->maps.push_back(m);
    return EXO_OK;
}

static int compile_and_upload(exo_problem* p)
{
    std::string err;
    Program prog;
    if (!compile_program(p->tree->model, p->maps, prog, err, p->compile_mode)) return fail(EXO_ERR_INVALID, err);
    p->prog = prog;
    if (p->d_image_bytes < p->prog.image.size())
    {
        if (p->d_image) cudaFree(p->d_image);
        p->d_image = nullptr;
        CUDA_TRY(cudaMalloc((void**)&p->d_image, p->prog.image.size()));
        p->d_image_bytes = p->prog.image.size();
    }
    // Kernels of earlier launches may still be staging / reading the old image on non-blocking streams (the library's own
    // pipeline streams, exo_stream_create): a legacy-stream cudaMemcpy does not order against those.  Recompiles are rare
    // (model state / joint limits changed), so drain the device before overwriting.
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(p->d_image, p->prog.image.data(), p->prog.image.size(), cudaMemcpyHostToDevice));
    p->chain_ok = chain_lower(p->prog, p->chain);
    p->plan_valid = false;
    p->compiled_version = p->tree->version;
    return EXO_OK;
}

int exo_problem_finalize(exo_problem* p)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (p->finalized) return fail(EXO_ERR_STATE, "problem already finalized");
    // host-only part first so that size queries work without a device
    std::string err;
    int lp = 0, lj = 0;
    if (!assign_indexing(p->tree->model, p->maps, lp, lj, err)) return fail(EXO_ERR_INVALID, err);
    Program prog;
    if (!compile_program(p->tree->model, p->maps, prog, err)) return fail(EXO_ERR_INVALID, err);
    p->prog = prog;
    p->chain_ok = chain_lower(p->prog, p->chain);
    p->finalized = true;
    p->compiled_version = -1;  // device upload happens lazily at the first update (needs a GPU)
    return EXO_OK;
}

int exo_problem_length_phi(const exo_problem* p)
{
    return (p && p->finalized) ? p->prog.hdr.length_phi : -1;
}

int exo_problem_length_jacobian(const exo_problem* p)
{
    return (p && p->finalized) ? p->prog.hdr.length_jac : -1;
}

int exo_problem_num_maps(const exo_problem* p)
{
    return p ? (int)p->maps.size() : -1;
}

int exo_problem_map_indexing(const exo_problem* p, int map, int32_t out[4])
{
    if (!p || !out) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (map < 0 || map >= (int)p->maps.size()) return fail(EXO_ERR_INVALID, "map index out of range");
    out[0] = p->maps[map].start;
    out[1] = p->maps[map].length;
    out[2] = p->maps[map].start_jacobian;
    out[3] = p->maps[map].length_jacobian;
    return EXO_OK;
}

int exo_problem_program_info(const exo_problem* p, int32_t out[8])
{
    if (!p || !out) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    out[0] = p->prog.n_joint_ops;
    out[1] = p->prog.n_emit_ops;
    out[2] = p->prog.hdr.n_frame_slots;
    out[3] = p->prog.hdr.n_joint_slots;
    out[4] = (int32_t)((p->prog.hdr.scratch_doubles + p->prog.hdr.n_ctrl) * sizeof(double));
    out[5] = (int32_t)p->prog.image.size();
    out[6] = p->prog.hdr.n_acc;
    out[7] = p->prog.hdr.n_ops;
    return EXO_OK;
}

int exo_problem_chain_info(const exo_problem* p, int32_t out[4])
{
    if (!p || !out) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    out[0] = p->chain_ok ? 1 : 0;
    out[1] = p->chain_ok ? p->chain.n_joints : 0;
    out[2] = p->chain_ok ? (p->chain.all_revolute | (p->chain.ortho << 1)) : 0;
    out[3] = p->chain_ok ? p->chain.n_frames : 0;
    return EXO_OK;
}

int exo_problem_set_generated(exo_problem* p, int32_t T, int32_t n_generated, const double* frames)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (T < 1) return fail(EXO_ERR_INVALID, "T must be >= 1");
    if (n_generated != p->tree->model.n_generated) return fail(EXO_ERR_INVALID, "n_generated does not match the tree's trajectory-driven elements");
    if (n_generated > 0 && !frames) return fail(EXO_ERR_INVALID, "frames is NULL");
    p->T = T;
    p->n_generated = n_generated;
    p->h_generated.assign(frames, frames + (size_t)T * n_generated * 12);
    if (p->d_generated)
    {
        cudaFree(p->d_generated);
        p->d_generated = nullptr;
    }
    if (n_generated > 0)
    {
        CUDA_TRY(cudaMalloc((void**)&p->d_generated, p->h_generated.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(p->d_generated, p->h_generated.data(), p->h_generated.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return EXO_OK;
}

static int fill_strides(const exo_problem* p, int layout, int64_t E, KernelParams& k)
{
    const int64_t N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, LJ = p->prog.hdr.length_jac;
    switch (layout)
    {
        case EXO_LAYOUT_SOA:
            k.x_se = 1; k.x_sk = E;
            k.phi_se = 1; k.phi_sk = E;
            k.j_se = 1; k.j_sr = N * E; k.j_sc = E;
            break;
        case EXO_LAYOUT_AOS_ROWMAJOR:
            k.x_se = N; k.x_sk = 1;
            k.phi_se = LP; k.phi_sk = 1;
            k.j_se = LJ * N; k.j_sr = N; k.j_sc = 1;
            break;
        case EXO_LAYOUT_AOS_COLMAJOR:
            k.x_se = N; k.x_sk = 1;
            k.phi_se = LP; k.phi_sk = 1;
            k.j_se = LJ * N; k.j_sr = 1; k.j_sc = LJ;
            break;
        default:
            return fail(EXO_ERR_INVALID, "unknown layout");
    }
    return EXO_OK;
}

// upper bound of one buffer of parked SphereCollision slots (EXO_SPH_DUMP_MIB, default 1024 MiB)
static int64_t sph_dump_cap_bytes()
{
    const char* s = getenv("EXO_SPH_DUMP_MIB");  // read per launch: tests shrink it to force several chunks
    const long mib = s ? atol(s) : 0;
    return (int64_t)(mib > 0 ? mib : 1024) << 20;
}

// Source of the per-robot kernel for the current model state (generated from the two-pass form of the program)
static bool tree_prepare(exo_problem* p)
{
    if (p->tree_version == p->tree->version) return p->tree_ok;
    p->tree_version = p->tree->version;
    p->tree_ok = false;
    p->tree_kernel = nullptr;
    p->tree_jit_failed = false;
    p->tree_key = exo_problem::TmaKey();
    std::string err;
    Program tp;
    std::vector<MapSpec> maps = p->maps;
    if (!compile_program(p->tree->model, maps, tp, err, COMPILE_TWO_PASS))
    {
        p->tree_why = err;
        return false;
    }
    p->tree_ok = tree_codegen(tp, p->tree_plan, p->tree_why);
    return p->tree_ok;
}

static int64_t tree_min_evals()
{
    static int64_t n = -1;
    if (n < 0)
    {
        const char* env = getenv("EXO_TREE_MIN_EVALS");
        n = env ? atoll(env) : 4096;
        if (n < 1) n = 1;
    }
    return n;
}

// e_offset: index of this launch's first evaluation in the caller's batch (the host pipeline launches chunk by chunk; per-evaluation
// side outputs such as the minimum sphere distance are indexed by the caller's evaluation number)
static int launch(exo_problem* p, const double* d_x, int64_t E, int64_t stride_E, int32_t T, int32_t layout, double* d_phi, double* d_jac,
                  cudaStream_t stream, bool timed, int64_t e_offset = 0)
{
    if (p->compiled_version != p->tree->version)
    {
        int rc = compile_and_upload(p);
        if (rc != EXO_OK) return rc;
    }
    if (p->prog.hdr.n_generated > 0)
    {
        if (!p->d_generated) return fail(EXO_ERR_STATE, "trajectory-driven elements need exo_problem_set_generated first");
        if (T != p->T) return fail(EXO_ERR_INVALID, "T does not match exo_problem_set_generated");
    }
    KernelParams k;
    memset(&k, 0, sizeof(k));
    k.program = p->d_image;
    k.program_bytes = (int32_t)p->prog.image.size();
    k.dbl_offset = p->prog.dbl_offset;
    k.T = T < 1 ? 1 : T;
    k.x = d_x;
    k.phi = d_phi;
    k.jac = d_jac;
    k.generated = p->d_generated;
    k.min_dist = p->d_min_dist ? p->d_min_dist + e_offset : nullptr;
    k.n_evals = E;
    int rc = fill_strides(p, layout, stride_E, k);
    if (rc != EXO_OK) return rc;
    const bool chain = p->chain_ok && p->kernel_choice != 1 && p->kernel_choice != 9;
    if (!chain && p->kernel_choice >= 2 && p->kernel_choice != 9) return fail(EXO_ERR_UNSUPPORTED, "the compiled program is not a serial chain: the chain kernel cannot run it");
    LaunchConfig cfg;
    int n_sph_ops = 0;
    for (const Op& op : p->prog.ops) n_sph_ops += op.kind == OP_EMIT_SPHERES;
    const bool split = !chain && n_sph_ops > 0 && n_sph_ops <= SPH_SPLIT_MAX_OPS && !p->fused_spheres;
    int64_t chunk = E;  // split launches go in chunks whose parked slots fit the dump buffer
    double* dump = nullptr;
    // tree programs: the per-robot generated kernel (TMA output) when the buffers qualify; run-time compiled on first use, so
    // small batches stay on the interpreter unless exo_problem_set_kernel(9) asks for it
    bool gen = false;
    // a solver that keeps calling with a smaller batch earns the compiler after a few calls (the interpreter serves the first ones)
    if (!chain && n_sph_ops == 0 && p->kernel_choice == 0 && E < tree_min_evals() && E >= 256 && p->tree_small_calls < 1000000) p->tree_small_calls++;
    const bool worth_compiling = E >= tree_min_evals() || (E >= 256 && p->tree_small_calls >= 16);
    if (!chain && n_sph_ops == 0 && (p->kernel_choice == 9 || (p->kernel_choice == 0 && worth_compiling && !p->tree_jit_failed)))
    {
        if (tree_prepare(p))
        {
            exo_problem::TmaKey& key = p->tree_key;
            if (key.phi != k.phi || key.jac != k.jac || key.n != E || key.phi_sk != k.phi_sk || key.j_sc != k.j_sc || key.j_sr != k.j_sr || key.version != p->tree_version ||
                p->tree_params.x != k.x || p->tree_params.x_sk != k.x_sk)
            {
                key.phi = k.phi; key.jac = k.jac; key.n = E; key.phi_sk = k.phi_sk; key.j_sc = k.j_sc; key.j_sr = k.j_sr; key.version = p->tree_version;
                key.ok = tree_jit_bind(p->tree_plan, p->tree_params, k.x, k.phi, k.jac, E, k.x_se, k.x_sk, k.phi_se, k.phi_sk, k.j_se, k.j_sr, k.j_sc);
                if (!key.ok) p->tree_why = "the generated kernel needs EXO_LAYOUT_SOA device buffers with an even, 16-byte aligned row pitch";
            }
            if (key.ok && !p->tree_kernel && !p->tree_jit_failed)
            {
                std::string jerr;
                p->tree_kernel = tree_jit_get(p->tree_plan, jerr);
                if (!p->tree_kernel)
                {
                    p->tree_jit_failed = true;
                    p->tree_why = "run-time compilation failed: " + jerr;
                }
            }
            gen = key.ok && p->tree_kernel;
        }
        if (!gen && p->kernel_choice == 9) return fail(EXO_ERR_UNSUPPORTED, "generated tree kernel not applicable: " + p->tree_why);
    }
    else if (p->kernel_choice == 9)
        return fail(EXO_ERR_UNSUPPORTED, "generated tree kernel not applicable: SphereCollision");
    if (!chain && !gen)
    {
        if (!p->plan_valid || p->plan_split != split)
        {
            cudaError_t e = plan_launch(p->prog.hdr, k.program_bytes, (int64_t)1 << 40, split, p->plan_cfg, k);
            if (e == cudaErrorInvalidConfiguration)
                return fail(EXO_ERR_UNSUPPORTED, "compiled program needs more shared memory per thread than one CTA can hold");
            if (e != cudaSuccess) return cuda_fail(e, "plan_launch");
            p->plan_scratch_offset = k.scratch_offset;
            p->plan_valid = true;
            p->plan_split = split;
        }
        cfg = p->plan_cfg;
        k.scratch_offset = p->plan_scratch_offset;
        if (split)
        {
            const int64_t row = (p->prog.hdr.scratch_doubles + 1) & ~1;
            const int64_t per_eval = (int64_t)n_sph_ops * row * (int64_t)sizeof(double);
            const int64_t cap_bytes = sph_dump_cap_bytes();
            int64_t max_chunk = cap_bytes / per_eval / 32 * 32;
            if (max_chunk < 32) max_chunk = 32;
            if (k.T > 1 && max_chunk > k.T) max_chunk = max_chunk / k.T * k.T;  // whole trajectories per chunk where they fit
            if (chunk > max_chunk) chunk = max_chunk;
            k.sph_row = (int32_t)row;
            k.sph_block = (chunk + 31) / 32 * 32 * row;  // whole tiles of 32 evaluations (the generated park kernel writes tiled rows)
            k.sph_n_ops = n_sph_ops;
            const int64_t need = (int64_t)n_sph_ops * k.sph_block * (int64_t)sizeof(double);
            exo_problem::SphDump* slot = nullptr;
            for (auto& sd : p->sph_dumps)
                if (sd.stream == stream) slot = &sd;
            if (!slot)
            {
                p->sph_dumps.push_back({stream, nullptr, 0});
                slot = &p->sph_dumps.back();
            }
            if (slot->bytes < need)
            {
                // the stream may still be reading the old buffer
                CUDA_TRY(cudaStreamSynchronize(stream));
                if (slot->d) CUDA_TRY(cudaFree(slot->d));
                slot->d = nullptr;
                slot->bytes = 0;
                CUDA_TRY(cudaMalloc(&slot->d, (size_t)need));
                slot->bytes = need;
            }
            dump = slot->d;
        }
    }
    if (timed)
    {
        if (!p->ev0)
        {
            CUDA_TRY(cudaEventCreate(&p->ev0));
            CUDA_TRY(cudaEventCreate(&p->ev1));
        }
        CUDA_TRY(cudaEventRecord(p->ev0, stream));
    }
    // front half of split SphereCollision programs: generated when the program holds nothing else (run-time compiled on first use)
    bool park = false;
    if (split && p->kernel_choice == 0 && E >= tree_min_evals() && !p->park_jit_failed)
    {
        if (p->park_version != p->compiled_version)
        {
            std::string why;
            p->park_version = p->compiled_version;
            p->park_kernel = nullptr;
            p->park_ok = tree_park_codegen(p->prog, p->park_plan, why);
        }
        if (p->park_ok && !p->park_kernel)
        {
            std::string jerr;
            p->park_kernel = tree_park_jit_get(p->park_plan, jerr);
            if (!p->park_kernel) p->park_jit_failed = true;
        }
        park = p->park_ok && p->park_kernel;
    }
    cudaError_t e;
    bool tma = false;
    if (chain)
    {
        chain_bind(p->chain, k.x, k.phi, k.jac, E, k.x_se, k.x_sk, k.phi_se, k.phi_sk, k.j_se, k.j_sr, k.j_sc);
        // SoA buffers of chains that own every column: outputs staged in shared memory and stored by the TMA unit
        if (p->kernel_choice == 0 || p->kernel_choice >= 7)
        {
            const int tile = p->kernel_choice == 7 ? 64 : (p->kernel_choice == 8 ? 128 : 0);
            exo_problem::TmaKey& key = p->chain_tma_key;
            if (key.phi != k.phi || key.jac != k.jac || key.n != E || key.phi_sk != k.phi_sk || key.j_sc != k.j_sc || key.j_sr != k.j_sr || key.tile != tile ||
                key.version != p->compiled_version)
            {
                key.phi = k.phi; key.jac = k.jac; key.n = E; key.phi_sk = k.phi_sk; key.j_sc = k.j_sc; key.j_sr = k.j_sr; key.tile = tile;
                key.version = p->compiled_version;
                key.ok = chain_tma_plan(p->chain, tile, p->chain_tma);
            }
            tma = key.ok;
            if (!tma && p->kernel_choice >= 7)
                return fail(EXO_ERR_UNSUPPORTED, "the TMA output path needs SoA buffers (even row pitch, 16-byte aligned) and a chain that owns every Jacobian column");
        }
        if (tma)
            e = chain_tma_launch(p->chain, p->chain_tma, stream);
        else
            e = chain_launch(p->chain, p->kernel_choice == 2 || p->kernel_choice == 3 ? p->kernel_choice - 2 : CHAIN_DEFAULT, stream);
        if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
        p->launches++;
    }
    else if (gen)
    {
        e = tree_jit_launch(p->tree_kernel, p->tree_plan, p->tree_params, stream);
        if (e != cudaSuccess) return cuda_fail(e, "generated kernel launch");
        p->launches++;
    }
    else
    {
        const KernelParams k0 = k;
        for (int64_t c0 = 0; c0 < E; c0 += chunk)
        {
            const int64_t n = E - c0 < chunk ? E - c0 : chunk;
            k.n_evals = n;
            k.e_base = c0;
            k.x = k0.x + c0 * k0.x_se;
            k.phi = k0.phi + c0 * k0.phi_se;
            k.jac = k0.jac ? k0.jac + c0 * k0.j_se : nullptr;
            k.min_dist = k0.min_dist ? k0.min_dist + c0 : nullptr;
            k.sph_dump = dump;
            k.sph_tiled = park ? 1 : 0;
            LaunchConfig c = cfg;
            const int64_t per_cta = (int64_t)(c.threads / 32) * c.tile;
            const int64_t want = (n + per_cta - 1) / per_cta;
            if (want < c.blocks) c.blocks = (int)want;
            if (park)
            {
                // split SphereCollision programs: the generated straight-line kernel parks the rows instead of the interpreter
                TreeParkParams pk;
                memset(&pk, 0, sizeof(pk));
                pk.x = k.x; pk.generated = k.generated; pk.dump = k.sph_dump;
                pk.n_evals = n; pk.x_se = k.x_se; pk.x_sk = k.x_sk; pk.e_base = k.e_base; pk.sph_block = k.sph_block;
                pk.T = k.T; pk.n_generated = p->prog.hdr.n_generated; pk.sph_row = k.sph_row;
                e = tree_park_jit_launch(p->park_kernel, p->park_plan, pk, stream);
            }
            else
                e = launch_eval(c, k, stream);
            if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
            p->launches++;
            if (split)
            {
                e = launch_spheres(c, k, stream);
                if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
                p->launches++;
            }
        }
    }
    p->last_kernel = gen ? "exo_tree_kernel" : chain ? (tma ? "exo_chain_tma_kernel" : "exo_chain_kernel") : (split ? "exo_spheres_kernel" : "exo_eval_kernel");
    if (timed)
    {
        CUDA_TRY(cudaEventRecord(p->ev1, stream));
        p->timed = true;
    }
    return EXO_OK;
}

int exo_problem_update_batch(exo_problem* p, const double* d_x, int64_t E, int32_t T, int32_t layout, double* d_phi, double* d_jac, void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (E < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (E == 0) return EXO_OK;
    if (!d_x || !d_phi) return fail(EXO_ERR_INVALID, "NULL buffer");
    return launch(p, d_x, E, E, T, layout, d_phi, d_jac, (cudaStream_t)stream, true);
}

int exo_problem_last_kernel_ms(exo_problem* p, float* ms)
{
    if (!p || !ms) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->timed) return fail(EXO_ERR_STATE, "no timed launch yet");
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return EXO_OK;
}

int exo_problem_set_kernel(exo_problem* p, int choice)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (choice < 0 || choice > 9)
        return fail(EXO_ERR_INVALID, "kernel choice must be 0 (auto), 1 (interpreter), 2..3 (chain kernel variants, per-lane stores), 4..5 (interpreter, one- / two-pass program), 6 (interpreter with SphereCollision fused in), 7..8 (chain kernel with TMA stores, 64 / 128 evaluations per tile) or 9 (per-robot generated tree kernel)");
    const int mode = choice == 4 ? COMPILE_ONE_PASS : (choice == 5 ? COMPILE_TWO_PASS : COMPILE_AUTO);
    if (mode != p->compile_mode)
    {
        p->compile_mode = mode;
        p->compiled_version = -1;  // recompile at the next launch
    }
    p->kernel_choice = (choice >= 4 && choice <= 6) ? 1 : choice;
    p->fused_spheres = choice == 6;
    return EXO_OK;
}

static bool tree_costgrad_prepare(exo_problem* p);
static bool update_cost_prepare(exo_problem* p);

int exo_problem_generated_kernel_source(exo_problem* p, char* out, int64_t capacity, int64_t* needed, int32_t compile)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    const bool cost_kernel = (compile & 2) != 0;  // bit 1: the fused cost + gradient kernel (needs exo_problem_set_cost) instead of the update kernel
    const bool uc_kernel = !cost_kernel && (compile & 4) != 0;  // bit 2: the update kernel with the cost epilogue (exo_problem_update_cost_batch)
    if (cost_kernel)
    {
        if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
        if (!tree_costgrad_prepare(p)) return fail(EXO_ERR_UNSUPPORTED, "generated cost kernel not applicable: " + p->tcg_why);
    }
    else if (uc_kernel)
    {
        if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
        if (!update_cost_prepare(p)) return fail(EXO_ERR_UNSUPPORTED, "generated update + cost kernel not applicable: " + p->uc_why);
    }
    else if (!tree_prepare(p))
        return fail(EXO_ERR_UNSUPPORTED, "generated tree kernel not applicable: " + p->tree_why);
    const std::string& src = cost_kernel ? p->tcg_plan.source : (uc_kernel ? p->uc_plan.source : p->tree_plan.source);
    if (needed) *needed = (int64_t)src.size() + 1;
    if (out && capacity > 0)
    {
        const size_t n = std::min((size_t)capacity - 1, src.size());
        memcpy(out, src.data(), n);
        out[n] = 0;
    }
    if (compile & 1)
    {
        std::vector<char> cubin;
        std::string err;
        const bool ok = cost_kernel ? tree_costgrad_jit_compile(p->tcg_plan, cubin, err) : tree_jit_compile(uc_kernel ? p->uc_plan : p->tree_plan, cubin, err);
        if (!ok) return fail(EXO_ERR_UNSUPPORTED, "run-time compilation failed: " + err);
    }
    return EXO_OK;
}

const char* exo_problem_last_kernel(const exo_problem* p)
{
    return p ? p->last_kernel : "none";
}

int64_t exo_problem_launch_count(const exo_problem* p)
{
    return p ? p->launches : -1;
}

static int ensure_pipeline(exo_problem* p, int64_t chunk, bool with_jac)
{
    if (p->pipe_chunk >= chunk && (p->pipe_has_jac || !with_jac)) return EXO_OK;
    free_pipeline(p);
    const size_t N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, LJ = p->prog.hdr.length_jac;
    for (int i = 0; i < exo_problem::NBUF; ++i)
    {
        if (!p->streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaMalloc((void**)&p->d_x[i], std::max<size_t>(8, chunk * N * sizeof(double))));
        CUDA_TRY(cudaMalloc((void**)&p->d_phi[i], std::max<size_t>(8, chunk * LP * sizeof(double))));
        if (with_jac) CUDA_TRY(cudaMalloc((void**)&p->d_jac[i], std::max<size_t>(8, chunk * LJ * N * sizeof(double))));
        CUDA_TRY(cudaMalloc((void**)&p->d_cost[i], std::max<size_t>(8, chunk * sizeof(double))));
        CUDA_TRY(cudaMalloc((void**)&p->d_grad[i], std::max<size_t>(8, chunk * N * sizeof(double))));
    }
    p->pipe_chunk = chunk;
    p->pipe_has_jac = with_jac;
    return EXO_OK;
}

int exo_problem_update_batch_host(exo_problem* p, const double* h_x, int64_t E, int32_t T, int32_t layout, double* h_phi, double* h_jac)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (E < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (E == 0) return EXO_OK;
    if (!h_x || !h_phi) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (layout < EXO_LAYOUT_SOA || layout > EXO_LAYOUT_AOS_COLMAJOR) return fail(EXO_ERR_INVALID, "unknown layout");
    const int64_t N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, LJ = p->prog.hdr.length_jac;
    // chunk: ~64 MiB of output per buffer (EXO_PIPE_CHUNK_MIB overrides; measured 4..128 MiB: 9.06, 8.60, 8.31, 8.15, 8.06, 8.12 ms per 461 MB step, PCIe D2H alone 7.75 ms), multiple of T so that e % T stays aligned per chunk
    static int64_t chunk_mib = 0;
    if (chunk_mib == 0)
    {
        const char* env = getenv("EXO_PIPE_CHUNK_MIB");
        chunk_mib = env ? atoll(env) : 64;
        if (chunk_mib < 1) chunk_mib = 64;
    }
    const int64_t bytes_per_eval = (N + LP + (h_jac ? LJ * N : 0)) * (int64_t)sizeof(double);
    int64_t chunk = std::max<int64_t>(1024, (chunk_mib << 20) / std::max<int64_t>(bytes_per_eval, 1));
    const int64_t Tm = T < 1 ? 1 : T;
    chunk = ((chunk + Tm - 1) / Tm) * Tm;
    if (chunk > E) chunk = E;
    int rc = ensure_pipeline(p, chunk, h_jac != nullptr);
    if (rc != EXO_OK) return rc;
    int b = 0;
    for (int64_t e0 = 0; e0 < E; e0 += chunk, b = (b + 1) % exo_problem::NBUF)
    {
        const int64_t n = std::min(chunk, E - e0);
        cudaStream_t s = p->streams[b];
        if (layout == EXO_LAYOUT_SOA)
        {
            // host arrays are [K][E]: copy the column block [K][e0:e0+n] into a dense [K][n] device array
            CUDA_TRY(cudaMemcpy2DAsync(p->d_x[b], n * sizeof(double), h_x + e0, E * sizeof(double), n * sizeof(double), N, cudaMemcpyHostToDevice, s));
        }
        else
            CUDA_TRY(cudaMemcpyAsync(p->d_x[b], h_x + e0 * N, n * N * sizeof(double), cudaMemcpyHostToDevice, s));
        rc = launch(p, p->d_x[b], n, n, T, layout, p->d_phi[b], h_jac ? p->d_jac[b] : nullptr, s, false, e0);
        if (rc != EXO_OK) return rc;
        if (layout == EXO_LAYOUT_SOA)
        {
            CUDA_TRY(cudaMemcpy2DAsync(h_phi + e0, E * sizeof(double), p->d_phi[b], n * sizeof(double), n * sizeof(double), LP, cudaMemcpyDeviceToHost, s));
            if (h_jac)
                CUDA_TRY(cudaMemcpy2DAsync(h_jac + e0, E * sizeof(double), p->d_jac[b], n * sizeof(double), n * sizeof(double), LJ * N, cudaMemcpyDeviceToHost, s));
        }
        else
        {
            CUDA_TRY(cudaMemcpyAsync(h_phi + e0 * LP, p->d_phi[b], n * LP * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (h_jac) CUDA_TRY(cudaMemcpyAsync(h_jac + e0 * LJ * N, p->d_jac[b], n * LJ * N * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
    }
    for (int i = 0; i < exo_problem::NBUF; ++i) CUDA_TRY(cudaStreamSynchronize(p->streams[i]));
    return EXO_OK;
}

// ---- auxiliary outputs, task cost, per-trajectory reductions ------------------------------------------------------------
int exo_problem_set_min_distance_output(exo_problem* p, double* d_min_distance)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (d_min_distance && p->prog.hdr.max_spheres == 0) return fail(EXO_ERR_INVALID, "the problem has no SphereCollision map");
    p->d_min_dist = d_min_distance;
    return EXO_OK;
}

static int upload(double** dst, const double* src, size_t n)
{
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    if (n == 0) return EXO_OK;
    CUDA_TRY(cudaMalloc((void**)dst, n * sizeof(double)));
    CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice));
    return EXO_OK;
}

int exo_problem_set_cost(exo_problem* p, const exo_cost_desc* d)
{
    if (!p || !d) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    const int n_maps = (int)p->maps.size();
    if (n_maps > COST_MAXMAPS) return fail(EXO_ERR_UNSUPPORTED, "more task maps than the cost kernel supports");
    if (d->T < 1) return fail(EXO_ERR_INVALID, "T must be >= 1");
    if (d->T_goal != 1 && d->T_goal != d->T) return fail(EXO_ERR_INVALID, "T_goal must be 1 or T");
    if (!d->rho) return fail(EXO_ERR_INVALID, "rho is NULL");
    const int LP = p->prog.hdr.length_phi, N = p->prog.hdr.n_ctrl;
    CostParams& c = p->cost;
    memset(&c, 0, sizeof(c));
    c.T = d->T;
    c.t_first = d->T > 1 ? 1 : 0;  // GetCost sums t = 1 .. T-1 (abstract_time_indexed_problem.cpp:299-307)
    c.T_goal = d->T_goal;
    c.n_maps = n_maps;
    c.n_ctrl = N;
    c.length_phi = LP;
    c.ct = d->ct;
    // the SetZero goal (task_space_vector.cpp:54-62) unless the caller gives one
    std::vector<double> goal((size_t)d->T_goal * LP, 0.0), goal_rot;
    int n_rot = 0;
    for (int k = 0; k < n_maps; ++k)
    {
        const MapSpec& m = p->maps[k];
        CostMapP& cm = c.map[k];
        cm.kind = m.kind;
        cm.start = m.start;
        cm.start_jacobian = m.start_jacobian;
        cm.length_jacobian = m.length_jacobian;
        cm.length = m.length;
        cm.n_frames = m.n_frames;
        cm.rot_type = m.rot_type;
        const bool lie = m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION;  // maps with SO(3) entries
        const int lead = m.kind == EXO_MAP_EFF_FRAME ? 3 : 0;
        cm.stride = lie ? lead + rotation_length(m.rot_type) : 0;
        if (m.kind == EXO_MAP_FRAMES) return fail(EXO_ERR_UNSUPPORTED, "raw-frame maps have no task cost");
        if (lie)
        {
            n_rot += m.n_frames;
            for (int t = 0; t < d->T_goal; ++t)
                for (int f = 0; f < m.n_frames; ++f)
                {
                    double* r = &goal[(size_t)t * LP + m.start + f * cm.stride + lead];
                    if (m.rot_type == EXO_ROT_QUATERNION) r[3] = 1.0;
                    if (m.rot_type == EXO_ROT_MATRIX) r[0] = r[4] = r[8] = 1.0;
                }
        }
    }
    if (d->goal) goal.assign(d->goal, d->goal + (size_t)d->T_goal * LP);
    // a binding that forwards TimeIndexedTask::y hands over one goal per timestep; when they are all the same row (the usual case: one
    // target for the whole trajectory) the kernels get ONE row, which every lane of a warp reads as a broadcast instead of 32 rows
    // 8 * length_Phi bytes apart (update + cost 83.8 -> 76.4 us on cfg3)
    int T_goal = d->T_goal;
    if (T_goal > 1)
    {
        bool same = true;
        for (int t = 1; t < T_goal && same; ++t) same = memcmp(&goal[(size_t)t * LP], &goal[0], sizeof(double) * (size_t)LP) == 0;
        if (same)
        {
            T_goal = 1;
            goal.resize((size_t)LP);
            c.T_goal = 1;
        }
    }
    c.n_rot = n_rot;
    goal_rot.assign((size_t)T_goal * n_rot * 9 + 9, 0.0);
    for (int t = 0; t < T_goal; ++t)
    {
        int ri = 0;
        for (int k = 0; k < n_maps; ++k)
        {
            const MapSpec& m = p->maps[k];
            if (m.kind != EXO_MAP_EFF_FRAME && m.kind != EXO_MAP_EFF_ORIENTATION) continue;
            const int lead = m.kind == EXO_MAP_EFF_FRAME ? 3 : 0;
            for (int f = 0; f < m.n_frames; ++f, ++ri)
                host_get_rotation(&goal[(size_t)t * LP + m.start + f * c.map[k].stride + lead], m.rot_type, &goal_rot[((size_t)t * n_rot + ri) * 9]);
        }
    }
    int rc = upload(&p->d_rho, d->rho, (size_t)d->T * n_maps);
    if (rc == EXO_OK) rc = upload(&p->d_goal, goal.data(), goal.size());
    if (rc == EXO_OK) rc = upload(&p->d_goal_rot, goal_rot.data(), goal_rot.size());
    if (rc == EXO_OK) rc = upload(&p->d_W, d->W, d->W ? (size_t)N : 0);
    if (rc != EXO_OK) return rc;
    p->cost_rho_t0.assign(d->rho, d->rho + n_maps);
    p->cost_rho.assign(d->rho, d->rho + (size_t)d->T * n_maps);
    p->cost_W.clear();
    if (d->W) p->cost_W.assign(d->W, d->W + N);
    c.rho = p->d_rho;
    c.goal = p->d_goal;
    c.goal_rot = p->d_goal_rot;
    c.W = p->d_W;
    p->cost_set = true;
    p->cost_epoch++;
    return EXO_OK;
}

int exo_problem_cost_batch(exo_problem* p, const double* d_x, const double* d_phi, int64_t n_traj, int32_t layout, double* d_cost_traj,
                           double* d_cost_eval, void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!d_x || !d_phi || (!d_cost_traj && !d_cost_eval)) return fail(EXO_ERR_INVALID, "NULL buffer");
    KernelParams k;
    memset(&k, 0, sizeof(k));
    int rc = fill_strides(p, layout, n_traj * p->cost.T, k);
    if (rc != EXO_OK) return rc;
    CostParams c = p->cost;
    c.x = d_x;
    c.phi = d_phi;
    c.cost_traj = d_cost_traj;
    c.cost_eval = d_cost_eval;
    c.n_traj = n_traj;
    c.x_se = k.x_se; c.x_sk = k.x_sk; c.phi_se = k.phi_se; c.phi_sk = k.phi_sk;
    cudaError_t e = launch_cost(c, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cost kernel launch");
    p->launches++;
    p->last_kernel = "exo_cost_kernel";
    return EXO_OK;
}

int exo_problem_cost_jacobian_batch(exo_problem* p, const double* d_x, const double* d_phi, const double* d_jac, int64_t n_traj, int32_t layout,
                                    double* d_grad, void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!d_x || !d_phi || !d_jac || !d_grad) return fail(EXO_ERR_INVALID, "NULL buffer");
    KernelParams k;
    memset(&k, 0, sizeof(k));
    const int64_t E = n_traj * p->cost.T;
    int rc = fill_strides(p, layout, E, k);
    if (rc != EXO_OK) return rc;
    CostGradParams g;
    memset(&g, 0, sizeof(g));
    g.c = p->cost;
    g.c.x = d_x;
    g.c.phi = d_phi;
    g.c.n_traj = n_traj;
    g.c.x_se = k.x_se; g.c.x_sk = k.x_sk; g.c.phi_se = k.phi_se; g.c.phi_sk = k.phi_sk;
    g.jac = d_jac;
    g.grad = d_grad;
    g.j_se = k.j_se; g.j_sr = k.j_sr; g.j_sc = k.j_sc;
    g.g_se = k.x_se; g.g_sk = k.x_sk;  // the gradient has the shape and layout of the states
    g.length_jac = p->prog.hdr.length_jac;
    cudaError_t e = launch_cost_gradient(g, (cudaStream_t)stream);
    if (e == cudaErrorInvalidConfiguration) return fail(EXO_ERR_UNSUPPORTED, "the task-space Jacobian has more rows than the cost-gradient kernel supports (200)");
    if (e != cudaSuccess) return cuda_fail(e, "cost gradient kernel launch");
    p->launches++;
    p->last_kernel = "exo_cost_gradient_kernel";
    return EXO_OK;
}

// Can exo_chain_costgrad_kernel serve this problem and cost?  Fills the per-frame map / goal-rotation indices if so.
static bool fused_costgrad_eligible(exo_problem* p, ChainGradParams& g)
{
    if (!p->chain_ok || p->kernel_choice == 1) return false;
    const ChainParams& P = p->chain;
    const CostParams& c = p->cost;
    std::vector<char> covered(p->maps.size(), 0);
    g.jl_map = -1;
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        if (fp.row_mode > ROWS_V) return false;
        int k_of = -1, rot = 0;
        for (int k = 0; k < c.n_maps && k_of < 0; ++k)
        {
            const CostMapP& m = c.map[k];
            const bool lie = m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION;
            if (fp.phi_off >= m.start && fp.phi_off < m.start + m.length)
            {
                if (m.kind != EXO_MAP_EFF_FRAME && m.kind != EXO_MAP_EFF_POSITION) return false;
                k_of = k;
                g.frame_rot[f] = m.kind == EXO_MAP_EFF_FRAME ? rot + (fp.phi_off - m.start) / m.stride : -1;
            }
            else if (lie)
                rot += m.n_frames;
        }
        if (k_of < 0) return false;
        g.frame_map[f] = k_of;
        covered[k_of] = 1;
    }
    for (int k = 0; k < c.n_maps; ++k)
    {
        if (covered[k]) continue;
        if (c.map[k].kind == EXO_MAP_JOINT_LIMIT && P.jl_enabled && c.map[k].start == P.jl_phi_off && g.jl_map < 0)
        {
            g.jl_map = k;
            continue;
        }
        // anything else (unused maps filled with their SetZero values) must not carry a cost
        for (int t = 0; t < c.T; ++t)
            if (p->cost_rho[(size_t)t * c.n_maps + k] != 0.0) return false;
    }
    return true;
}

// Can the generated cost kernel serve this problem and cost?  Fills the per-frame cost map / goal rotation indices (position slots
// follow the order of the frame emits in the two-pass program, which is the order of the maps' frames).
static bool tree_costgrad_spec(exo_problem* p, const Program& tp, TreeCostSpec& spec)
{
    const CostParams& c = p->cost;
    spec = TreeCostSpec();
    std::vector<char> covered(p->maps.size(), 0);
    for (const Op& op : tp.ops)
    {
        if (op.kind == OP_END) break;
        if (op.kind != OP_EMIT_FRAME) continue;
        int k_of = -1, rot = 0, frame_rot = -1;
        for (int k = 0; k < c.n_maps && k_of < 0; ++k)
        {
            const CostMapP& m = c.map[k];
            const bool lie = m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION;
            if (op.phi_off >= m.start && op.phi_off < m.start + m.length)
            {
                if (m.kind != EXO_MAP_EFF_FRAME && m.kind != EXO_MAP_EFF_POSITION) return false;
                k_of = k;
                frame_rot = m.kind == EXO_MAP_EFF_FRAME ? rot + (op.phi_off - m.start) / m.stride : -1;
            }
            else if (lie)
                rot += m.n_frames;
        }
        if (k_of < 0) return false;
        if ((int)spec.frame_map.size() <= op.n_items)
        {
            spec.frame_map.resize(op.n_items + 1, -1);
            spec.frame_rot.resize(op.n_items + 1, -1);
        }
        spec.frame_map[op.n_items] = k_of;
        spec.frame_rot[op.n_items] = frame_rot;
        covered[k_of] = 1;
    }
    for (int k = 0; k < c.n_maps; ++k)
    {
        if (covered[k]) continue;
        if (c.map[k].kind == EXO_MAP_JOINT_LIMIT && spec.jl_map < 0)
        {
            spec.jl_map = k;
            continue;
        }
        // anything else (unused maps filled with their SetZero values, maps without frames) must not carry a cost
        for (int t = 0; t < c.T; ++t)
            if (p->cost_rho[(size_t)t * c.n_maps + k] != 0.0) return false;
    }
    return true;
}

// Source of the generated cost kernel for the current model state and cost structure
static bool tree_costgrad_prepare(exo_problem* p)
{
    if (p->chain_ok || p->kernel_choice == 1) return false;
    if (p->tcg_version == p->tree->version && p->tcg_cost_epoch == p->cost_epoch) return p->tcg_ok;
    p->tcg_cost_epoch = p->cost_epoch;
    std::string err;
    Program tp;
    std::vector<MapSpec> maps = p->maps;
    if (!compile_program(p->tree->model, maps, tp, err, COMPILE_TWO_PASS))
    {
        p->tcg_why = err;
        p->tcg_version = p->tree->version;
        return p->tcg_ok = false;
    }
    TreeCostSpec spec;
    if (!tree_costgrad_spec(p, tp, spec))
    {
        p->tcg_why = "the cost is not made of EffFrame / EffPosition / JointLimit maps";
        p->tcg_version = p->tree->version;
        return p->tcg_ok = false;
    }
    const bool same = p->tcg_version == p->tree->version && spec.frame_map == p->tcg_spec.frame_map && spec.frame_rot == p->tcg_spec.frame_rot &&
                      spec.jl_map == p->tcg_spec.jl_map;
    if (same) return p->tcg_ok;
    p->tcg_version = p->tree->version;
    p->tcg_spec = spec;
    p->tcg_kernel = nullptr;
    p->tcg_jit_failed = false;
    p->tcg_ok = tree_costgrad_codegen(tp, spec, p->tcg_plan, p->tcg_why);
    return p->tcg_ok;
}

// scratch / scratch_bytes: the buffer holding per-evaluation costs (+ Phi, J when the program is not fused); the public entry point
// uses the problem's own, the host pipeline one per stream so that its chunks overlap
static int cost_gradient_impl(exo_problem* p, const double* d_x, int64_t n_traj, int32_t layout, double* d_cost_traj, double* d_grad, void* stream,
                              double** scratch, int64_t* scratch_bytes)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!d_x || (!d_grad && !d_cost_traj)) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (p->compiled_version != p->tree->version)
    {
        int rc = compile_and_upload(p);
        if (rc != EXO_OK) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t T = p->cost.T, E = n_traj * T;
    const int64_t N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, LJ = p->prog.hdr.length_jac;
    if (p->prog.hdr.n_generated > 0 && T != p->T) return fail(EXO_ERR_INVALID, "T of the cost does not match exo_problem_set_generated");
    KernelParams k;
    memset(&k, 0, sizeof(k));
    int rc = fill_strides(p, layout, E, k);
    if (rc != EXO_OK) return rc;
    ChainGradParams g;
    memset(&g, 0, sizeof(g));
    const bool fused = fused_costgrad_eligible(p, g);
    // tree programs: the generated fused kernel (run-time compiled on first use)
    bool tree_fused = false;
    if (!fused && !p->tcg_jit_failed && tree_costgrad_prepare(p))
    {
        if (!p->tcg_kernel)
        {
            std::string jerr;
            p->tcg_kernel = tree_costgrad_jit_get(p->tcg_plan, jerr);
            if (!p->tcg_kernel)
            {
                p->tcg_jit_failed = true;
                p->tcg_why = "run-time compilation failed: " + jerr;
            }
        }
        tree_fused = p->tcg_kernel != nullptr;
    }
    const int64_t need = (E + ((fused || tree_fused) ? 0 : E * (LP + LJ * N))) * (int64_t)sizeof(double);
    if (*scratch_bytes < need)
    {
        CUDA_TRY(cudaStreamSynchronize(s));
        if (*scratch) CUDA_TRY(cudaFree(*scratch));
        *scratch = nullptr;
        *scratch_bytes = 0;
        CUDA_TRY(cudaMalloc((void**)scratch, (size_t)need));
        *scratch_bytes = need;
    }
    double* d_cost_eval = *scratch;
    if (fused)
    {
        const CostParams& c = p->cost;
        g.rho = c.rho; g.goal = c.goal; g.goal_rot = c.goal_rot; g.W = c.W;
        g.grad = d_grad;
        g.cost_eval = d_cost_traj ? d_cost_eval : nullptr;
        g.ct = c.ct;
        g.g_se = k.x_se; g.g_sk = k.x_sk;
        g.T = c.T; g.t_first = c.t_first; g.T_goal = c.T_goal; g.n_maps = c.n_maps; g.length_phi = c.length_phi; g.n_rot = c.n_rot;
        chain_bind(p->chain, d_x, nullptr, nullptr, E, k.x_se, k.x_sk, k.phi_se, k.phi_sk, k.j_se, k.j_sr, k.j_sc);
        cudaError_t e = d_grad ? chain_costgrad_launch(p->chain, g, s) : chain_cost_launch(p->chain, g, s);  // d_grad NULL: cost only
        if (e != cudaSuccess) return cuda_fail(e, "fused cost-gradient kernel launch");
        p->launches++;
        if (d_cost_traj)
        {
            e = launch_reduce(d_cost_eval, n_traj, (int)T, 0, 1, d_cost_traj, s);
            if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
            p->launches++;
        }
        p->last_kernel = d_grad ? "exo_chain_costgrad_kernel" : "exo_chain_cost_kernel";
        return EXO_OK;
    }
    if (tree_fused)
    {
        const CostParams& c = p->cost;
        TreeCostParams tc;
        memset(&tc, 0, sizeof(tc));
        tc.x = d_x; tc.grad = d_grad; tc.cost_eval = d_cost_traj ? d_cost_eval : nullptr;
        tc.rho = c.rho; tc.goal = c.goal; tc.goal_rot = c.goal_rot; tc.W = c.W;
        tc.n_evals = E; tc.x_se = k.x_se; tc.x_sk = k.x_sk; tc.g_se = k.x_se; tc.g_sk = k.x_sk;
        tc.ct = c.ct;
        tc.T = c.T; tc.t_first = c.t_first; tc.T_goal = c.T_goal; tc.n_maps = c.n_maps; tc.length_phi = c.length_phi; tc.n_rot = c.n_rot;
        cudaError_t e = tree_costgrad_jit_launch(p->tcg_kernel, p->tcg_plan, tc, s);
        if (e != cudaSuccess) return cuda_fail(e, "generated cost-gradient kernel launch");
        p->launches++;
        if (d_cost_traj)
        {
            e = launch_reduce(d_cost_eval, n_traj, (int)T, 0, 1, d_cost_traj, s);
            if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
            p->launches++;
        }
        p->last_kernel = "exo_tree_costgrad_kernel";
        return EXO_OK;
    }
    // general programs: Phi and J into the scratch buffer, then the cost and cost-gradient kernels
    double* d_phi = *scratch + E;
    double* d_jac = d_grad ? d_phi + E * LP : nullptr;  // cost only: forward kinematics without Jacobians
    rc = launch(p, d_x, E, E, (int32_t)T, layout, d_phi, d_jac, s, false);
    if (rc != EXO_OK) return rc;
    if (d_cost_traj)
    {
        rc = exo_problem_cost_batch(p, d_x, d_phi, n_traj, layout, d_cost_traj, nullptr, stream);
        if (rc != EXO_OK) return rc;
    }
    if (!d_grad) return EXO_OK;
    return exo_problem_cost_jacobian_batch(p, d_x, d_phi, d_jac, n_traj, layout, d_grad, stream);
}

int exo_problem_cost_gradient_batch(exo_problem* p, const double* d_x, int64_t n_traj, int32_t layout, double* d_cost_traj, double* d_grad, void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    return cost_gradient_impl(p, d_x, n_traj, layout, d_cost_traj, d_grad, stream, &p->d_cg, &p->cg_bytes);
}

// Source of the generated update + cost kernel for the current model state and cost structure
static bool update_cost_prepare(exo_problem* p)
{
    if (p->uc_version == p->tree->version && p->uc_cost_epoch == p->cost_epoch) return p->uc_ok;
    p->uc_cost_epoch = p->cost_epoch;
    std::string err;
    Program tp;
    std::vector<MapSpec> maps = p->maps;
    if (!compile_program(p->tree->model, maps, tp, err, COMPILE_TWO_PASS))
    {
        p->uc_why = err;
        p->uc_version = p->tree->version;
        return p->uc_ok = false;
    }
    TreeCostSpec spec;
    if (!tree_costgrad_spec(p, tp, spec))
    {
        p->uc_why = "the cost is not made of EffFrame / EffPosition / JointLimit maps";
        p->uc_version = p->tree->version;
        return p->uc_ok = false;
    }
    const bool same = p->uc_version == p->tree->version && spec.frame_map == p->uc_spec.frame_map && spec.frame_rot == p->uc_spec.frame_rot &&
                      spec.jl_map == p->uc_spec.jl_map;
    if (same) return p->uc_ok;
    p->uc_version = p->tree->version;
    p->uc_spec = spec;
    p->uc_kernel = nullptr;
    p->uc_jit_failed = false;
    p->uc_key = exo_problem::TmaKey();
    p->uc_ok = tree_codegen(tp, p->uc_plan, p->uc_why, &p->uc_spec);
    return p->uc_ok;
}

int exo_problem_update_cost_batch(exo_problem* p, const double* d_x, int64_t n_traj, int32_t layout, double* d_phi, double* d_jac, double* d_cost_traj,
                                  double* d_cost_eval, void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!d_x || !d_phi || (!d_cost_traj && !d_cost_eval)) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (p->compiled_version != p->tree->version)
    {
        int rc = compile_and_upload(p);
        if (rc != EXO_OK) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t T = p->cost.T, E = n_traj * T;
    if (p->prog.hdr.n_generated > 0 && T != p->T) return fail(EXO_ERR_INVALID, "T of the cost does not match exo_problem_set_generated");
    // one generated kernel (forward kinematics, Jacobians and the cost of every evaluation while its Phi is in registers) + the
    // per-trajectory sum, for SoA buffers and costs made of EffFrame / EffPosition / JointLimit maps; run-time compiled, so small
    // batches and everything else take the two calls below
    const bool want = p->kernel_choice == 0 || p->kernel_choice == 9;
    if (want && layout == EXO_LAYOUT_SOA && !p->uc_jit_failed && (E >= tree_min_evals() || p->kernel_choice == 9) && update_cost_prepare(p))
    {
        KernelParams k;
        memset(&k, 0, sizeof(k));
        k.x = d_x; k.phi = d_phi; k.jac = d_jac; k.n_evals = E;
        int rc = fill_strides(p, layout, E, k);
        if (rc != EXO_OK) return rc;
        exo_problem::TmaKey& key = p->uc_key;
        if (key.phi != k.phi || key.jac != k.jac || key.n != E || key.phi_sk != k.phi_sk || key.j_sc != k.j_sc || key.j_sr != k.j_sr || key.version != p->uc_version ||
            p->uc_params.x != k.x || p->uc_params.x_sk != k.x_sk)
        {
            key.phi = k.phi; key.jac = k.jac; key.n = E; key.phi_sk = k.phi_sk; key.j_sc = k.j_sc; key.j_sr = k.j_sr; key.version = p->uc_version;
            key.ok = tree_jit_bind(p->uc_plan, p->uc_params, k.x, k.phi, k.jac, E, k.x_se, k.x_sk, k.phi_se, k.phi_sk, k.j_se, k.j_sr, k.j_sc);
        }
        if (key.ok && !p->uc_kernel)
        {
            std::string jerr;
            p->uc_kernel = tree_jit_get(p->uc_plan, jerr);
            if (!p->uc_kernel)
            {
                p->uc_jit_failed = true;
                p->uc_why = "run-time compilation failed: " + jerr;
            }
        }
        if (key.ok && p->uc_kernel)
        {
            double* ce = d_cost_eval;
            if (!ce)
            {
                const int64_t need = E * (int64_t)sizeof(double);
                if (p->uc_cost_bytes < need)
                {
                    CUDA_TRY(cudaStreamSynchronize(s));
                    if (p->d_uc_cost) CUDA_TRY(cudaFree(p->d_uc_cost));
                    p->d_uc_cost = nullptr;
                    p->uc_cost_bytes = 0;
                    CUDA_TRY(cudaMalloc((void**)&p->d_uc_cost, (size_t)need));
                    p->uc_cost_bytes = need;
                }
                ce = p->d_uc_cost;
            }
            const CostParams& c = p->cost;
            TreeGenParams& tp = p->uc_params;
            tp.cost_eval = ce;
            tp.rho = c.rho; tp.goal = c.goal; tp.goal_rot = c.goal_rot; tp.W = c.W;
            tp.ct = c.ct;
            tp.T = c.T; tp.t_first = c.t_first; tp.T_goal = c.T_goal; tp.n_maps = c.n_maps; tp.length_phi = c.length_phi; tp.n_rot = c.n_rot;
            cudaError_t e = tree_jit_launch(p->uc_kernel, p->uc_plan, tp, s);
            if (e != cudaSuccess) return cuda_fail(e, "generated update + cost kernel launch");
            p->launches++;
            if (d_cost_traj)
            {
                e = launch_reduce(ce, n_traj, (int)T, 0, 1, d_cost_traj, s);
                if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
                p->launches++;
            }
            p->last_kernel = "exo_tree_update_cost_kernel";
            return EXO_OK;
        }
    }
    int rc = launch(p, d_x, E, E, (int32_t)T, layout, d_phi, d_jac, s, false);
    if (rc != EXO_OK) return rc;
    return exo_problem_cost_batch(p, d_x, d_phi, n_traj, layout, d_cost_traj, d_cost_eval, stream);
}

int exo_problem_aico_task_messages_batch(exo_problem* p, const double* d_qhat, const double* d_phi, const double* d_jac, int64_t n_traj, int32_t layout,
                                         double* d_R, double* d_r, double* d_rhat, double* d_cost, void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!d_qhat || !d_phi || !d_jac || !d_R || !d_r || !d_rhat || !d_cost) return fail(EXO_ERR_INVALID, "NULL buffer");
    KernelParams k;
    memset(&k, 0, sizeof(k));
    const int64_t E = n_traj * p->cost.T, N = p->prog.hdr.n_ctrl;
    int rc = fill_strides(p, layout, E, k);
    if (rc != EXO_OK) return rc;
    AicoParams a;
    memset(&a, 0, sizeof(a));
    a.c = p->cost;
    a.c.x = d_qhat;
    a.c.phi = d_phi;
    a.c.n_traj = n_traj;
    a.c.x_se = k.x_se; a.c.x_sk = k.x_sk; a.c.phi_se = k.phi_se; a.c.phi_sk = k.phi_sk;
    a.jac = d_jac;
    a.j_se = k.j_se; a.j_sr = k.j_sr; a.j_sc = k.j_sc;
    a.R = d_R; a.r = d_r; a.rhat = d_rhat; a.cost = d_cost;
    a.r_se = k.x_se; a.r_sk = k.x_sk;
    if (layout == EXO_LAYOUT_SOA)
    {
        a.R_se = 1; a.R_si = N * E; a.R_sj = E;  // R[N][N][E]
    }
    else
    {
        a.R_se = N * N; a.R_si = N; a.R_sj = 1;  // R[E][N][N] (symmetric: row- and column-major coincide)
    }
    a.length_jac = p->prog.hdr.length_jac;
    cudaError_t e = launch_aico_messages(a, (cudaStream_t)stream);
    if (e == cudaErrorInvalidConfiguration) return fail(EXO_ERR_UNSUPPORTED, "the task-space Jacobian has more rows than the AICO message kernel supports (100)");
    if (e != cudaSuccess) return cuda_fail(e, "AICO message kernel launch");
    p->launches++;
    p->last_kernel = "exo_aico_messages_kernel";
    return EXO_OK;
}

int exo_trajectory_reduce(const double* d_in, int64_t n_traj, int32_t T, int32_t t_first, int32_t op, double* d_out, void* stream)
{
    if (n_traj < 0 || T < 1 || t_first < 0 || t_first >= T) return fail(EXO_ERR_INVALID, "bad trajectory shape");
    if (op != EXO_REDUCE_MIN && op != EXO_REDUCE_SUM) return fail(EXO_ERR_INVALID, "unknown reduction");
    if (n_traj == 0) return EXO_OK;
    if (!d_in || !d_out) return fail(EXO_ERR_INVALID, "NULL buffer");
    cudaError_t e = launch_reduce(d_in, n_traj, T, t_first, op, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
    return EXO_OK;
}

// ---- kernel + collective in one call (row e) ---------------------------------------------------------------------------------
int exo_problem_cost_allgather(exo_problem* p, exo_comm* comm, const double* d_x, int64_t n_traj_local, int64_t n_traj_per_rank, int32_t layout,
                               double* d_cost_all, void* stream)
{
    if (!p || !comm) return fail(EXO_ERR_INVALID, "NULL argument");
    if (n_traj_local < 0 || n_traj_per_rank < n_traj_local) return fail(EXO_ERR_INVALID, "a rank holds at most n_traj_per_rank trajectories");
    if (n_traj_per_rank == 0) return EXO_OK;
    if (!d_cost_all) return fail(EXO_ERR_INVALID, "NULL buffer");
    double* mine = d_cost_all + (int64_t)exo_comm_rank(comm) * n_traj_per_rank;
    if (n_traj_local > 0)
    {
        int rc = exo_problem_cost_gradient_batch(p, d_x, n_traj_local, layout, mine, nullptr, stream);
        if (rc != EXO_OK) return rc;
    }
    if (n_traj_local < n_traj_per_rank)
        CUDA_TRY(cudaMemsetAsync(mine + n_traj_local, 0, (size_t)(n_traj_per_rank - n_traj_local) * sizeof(double), (cudaStream_t)stream));
    return exo_comm_allgather(comm, mine, d_cost_all, n_traj_per_rank, stream);
}

extern "C" int exo_internal_min_gather(exo_comm* c, const double* d_min_traj_local, int64_t n_traj_local, int64_t n_traj_per_rank, double* d_min_traj_all,
                                       double* d_min_global, void* stream);

int exo_min_distance_allgather(exo_comm* comm, const double* d_min_eval, int64_t n_traj_local, int32_t T, int64_t n_traj_per_rank, double* d_min_traj_all,
                               double* d_min_global, void* stream)
{
    if (!comm) return fail(EXO_ERR_INVALID, "NULL argument");
    if (T < 1 || n_traj_local < 0 || n_traj_per_rank < n_traj_local) return fail(EXO_ERR_INVALID, "a rank holds at most n_traj_per_rank trajectories of T >= 1 timesteps");
    if (!d_min_traj_all && !d_min_global) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (n_traj_local > 0 && !d_min_eval) return fail(EXO_ERR_INVALID, "NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (d_min_traj_all && n_traj_per_rank > 0)
    {
        double* mine = d_min_traj_all + (int64_t)exo_comm_rank(comm) * n_traj_per_rank;
        e = launch_reduce(d_min_eval, n_traj_local, T, 0, EXO_REDUCE_MIN, mine, s);
        if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
        // several ranks with peer windows: padding, exchange and global minimum in one kernel over NVLink peer memory
        if (exo_comm_world_size(comm) > 1 && exo_internal_min_gather(comm, mine, n_traj_local, n_traj_per_rank, d_min_traj_all, d_min_global, stream) == EXO_OK)
            return EXO_OK;
        e = launch_fill(mine + n_traj_local, n_traj_per_rank - n_traj_local, (double)INFINITY, s);
        if (e != cudaSuccess) return cuda_fail(e, "fill kernel launch");
        int rc = exo_comm_allgather(comm, mine, d_min_traj_all, n_traj_per_rank, stream);
        if (rc != EXO_OK) return rc;
        if (d_min_global)
        {
            // every rank holds every trajectory's minimum now: the global one is a local reduction, no second collective
            e = launch_reduce(d_min_traj_all, 1, (int)std::min<int64_t>((int64_t)exo_comm_world_size(comm) * n_traj_per_rank, INT32_MAX), 0, EXO_REDUCE_MIN,
                              d_min_global, s);
            if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
        }
        return EXO_OK;
    }
    // global minimum only: local min over all evaluations, then all-reduce(min) of one double
    if (n_traj_local * (int64_t)T > INT32_MAX) return fail(EXO_ERR_UNSUPPORTED, "more than 2^31 evaluations per rank");
    if (n_traj_local > 0)
        e = launch_reduce(d_min_eval, 1, (int)(n_traj_local * T), 0, EXO_REDUCE_MIN, d_min_global, s);
    else
        e = launch_fill(d_min_global, 1, (double)INFINITY, s);
    if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
    return exo_comm_allreduce(comm, d_min_global, d_min_global, 1, EXO_REDUCE_MIN, stream);
}

namespace
{
struct EventGuard  // a CUDA event destroyed on every return path
{
    cudaEvent_t ev = nullptr;
    ~EventGuard()
    {
        if (ev) cudaEventDestroy(ev);
    }
};
}  // namespace

int exo_problem_cost_batch_host(exo_problem* p, const double* h_x, int64_t n_traj, int32_t layout, double* h_cost_traj)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!h_x || !h_cost_traj) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (layout != EXO_LAYOUT_AOS_ROWMAJOR && layout != EXO_LAYOUT_AOS_COLMAJOR)
        return fail(EXO_ERR_UNSUPPORTED, "host cost batches take evaluation-major states X[B][T][N]");
    const int64_t N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, T = p->cost.T;
    // whole trajectories per chunk, ~16 MiB of Phi per buffer; forward kinematics only (no Jacobian is produced or moved)
    int64_t traj_per_chunk = std::max<int64_t>(1, (16ll << 20) / std::max<int64_t>(1, (N + LP) * T * (int64_t)sizeof(double)));
    if (traj_per_chunk > n_traj) traj_per_chunk = n_traj;
    int rc = ensure_pipeline(p, traj_per_chunk * T, false);
    if (rc != EXO_OK) return rc;
    if (p->compiled_version != p->tree->version)
    {
        rc = compile_and_upload(p);
        if (rc != EXO_OK) return rc;
    }
    EventGuard prev_done;
    int b = 0;
    for (int64_t t0 = 0; t0 < n_traj; t0 += traj_per_chunk, b = (b + 1) % exo_problem::NBUF)
    {
        const int64_t nt = std::min(traj_per_chunk, n_traj - t0), n = nt * T;
        cudaStream_t s = p->streams[b];
        CUDA_TRY(cudaMemcpyAsync(p->d_x[b], h_x + t0 * T * N, n * N * sizeof(double), cudaMemcpyHostToDevice, s));
        ChainGradParams probe;
        memset(&probe, 0, sizeof(probe));
        if (fused_costgrad_eligible(p, probe))
        {
            // chain programs: one kernel, forward kinematics and cost in registers (its per-evaluation scratch is shared by the
            // chunks: serialise the kernels through an event, the copies still overlap)
            if (prev_done.ev) CUDA_TRY(cudaStreamWaitEvent(s, prev_done.ev, 0));
            rc = exo_problem_cost_gradient_batch(p, p->d_x[b], nt, layout, p->d_cost[b], nullptr, s);
            if (rc != EXO_OK) return rc;
            if (!prev_done.ev) CUDA_TRY(cudaEventCreateWithFlags(&prev_done.ev, cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(prev_done.ev, s));
        }
        else
        {
            rc = launch(p, p->d_x[b], n, n, (int32_t)T, layout, p->d_phi[b], nullptr, s, false);
            if (rc != EXO_OK) return rc;
            rc = exo_problem_cost_batch(p, p->d_x[b], p->d_phi[b], nt, layout, p->d_cost[b], nullptr, s);
            if (rc != EXO_OK) return rc;
        }
        CUDA_TRY(cudaMemcpyAsync(h_cost_traj + t0, p->d_cost[b], nt * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < exo_problem::NBUF; ++i) CUDA_TRY(cudaStreamSynchronize(p->streams[i]));
    return EXO_OK;
}

int exo_problem_cost_jacobian_batch_host(exo_problem* p, const double* h_x, int64_t n_traj, int32_t layout, double* h_cost_traj, double* h_grad)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (n_traj < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_traj == 0) return EXO_OK;
    if (!h_x || !h_grad) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (layout != EXO_LAYOUT_AOS_ROWMAJOR && layout != EXO_LAYOUT_AOS_COLMAJOR)
        return fail(EXO_ERR_UNSUPPORTED, "host cost batches take evaluation-major states X[B][T][N]");
    const int64_t N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, LJ = p->prog.hdr.length_jac, T = p->cost.T;
    // whole trajectories per chunk, ~32 MiB of device-side Phi + J per buffer; Phi and J never leave the device: N doubles per
    // evaluation (and one cost per trajectory) come back instead of length_jacobian x N
    int64_t traj_per_chunk = std::max<int64_t>(1, (32ll << 20) / std::max<int64_t>(1, (N + LP + LJ * N) * T * (int64_t)sizeof(double)));
    {
        // fused programs keep nothing but the states and the gradient on the device: chunks of ~4 MiB of states (large enough for
        // the link, small enough that the first copy back starts early), one eighth of the batch at most
        if (p->compiled_version != p->tree->version)
        {
            const int rcc = compile_and_upload(p);
            if (rcc != EXO_OK) return rcc;
        }
        ChainGradParams probe;
        memset(&probe, 0, sizeof(probe));
        if (fused_costgrad_eligible(p, probe))
        {
            traj_per_chunk = std::max<int64_t>(1, (4ll << 20) / std::max<int64_t>(1, N * T * (int64_t)sizeof(double)));
            traj_per_chunk = std::min(traj_per_chunk, std::max<int64_t>(1, (n_traj + exo_problem::NBUF - 1) / exo_problem::NBUF));
        }
    }
    if (traj_per_chunk > n_traj) traj_per_chunk = n_traj;
    int rc = ensure_pipeline(p, traj_per_chunk * T, false);
    if (rc != EXO_OK) return rc;
    int b = 0;
    for (int64_t t0 = 0; t0 < n_traj; t0 += traj_per_chunk, b = (b + 1) % exo_problem::NBUF)
    {
        const int64_t nt = std::min(traj_per_chunk, n_traj - t0), n = nt * T;
        cudaStream_t s = p->streams[b];
        CUDA_TRY(cudaMemcpyAsync(p->d_x[b], h_x + t0 * T * N, n * N * sizeof(double), cudaMemcpyHostToDevice, s));
        // every stream has its own scratch, so the chunks' kernels and copies overlap freely
        rc = cost_gradient_impl(p, p->d_x[b], nt, layout, h_cost_traj ? p->d_cost[b] : nullptr, p->d_grad[b], s, &p->d_cgs[b], &p->cgs_bytes[b]);
        if (rc != EXO_OK) return rc;
        if (h_cost_traj) CUDA_TRY(cudaMemcpyAsync(h_cost_traj + t0, p->d_cost[b], nt * sizeof(double), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(h_grad + t0 * T * N, p->d_grad[b], n * N * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < exo_problem::NBUF; ++i) CUDA_TRY(cudaStreamSynchronize(p->streams[i]));
    return EXO_OK;
}

// ---- task-map Hessians ------------------------------------------------------------------------------------------------------
static int ensure_shadow(exo_problem* p)
{
    if (p->shadow) return EXO_OK;
    int rc = exo_problem_create(p->tree, &p->shadow);
    if (rc != EXO_OK) return rc;
    std::vector<int32_t> ea, eb;
    std::vector<double> oa, ob;
    p->shadow_frame_start.clear();
    for (const MapSpec& m : p->maps)
    {
        p->shadow_frame_start.push_back((int)ea.size());
        if (m.kind == EXO_MAP_SPHERE_COLLISION || m.kind == EXO_MAP_JOINT_LIMIT || m.kind == EXO_MAP_JOINT_POSE ||
            m.kind == EXO_MAP_EFF_AXIS_ALIGNMENT || m.kind == EXO_MAP_CENTER_OF_MASS)
            continue;  // no Hessian overload in the reference: the base-class fallback needs the map's own Jacobian only
        ea.insert(ea.end(), m.elem_a.begin(), m.elem_a.end());
        eb.insert(eb.end(), m.elem_b.begin(), m.elem_b.end());
        oa.insert(oa.end(), m.off_a.begin(), m.off_a.end());
        ob.insert(ob.end(), m.off_b.begin(), m.off_b.end());
    }
    if (!ea.empty())
    {
        exo_task_map_desc d;
        memset(&d, 0, sizeof(d));
        d.kind = EXO_MAP_FRAMES;
        d.n_frames = (int)ea.size();
        d.elem_a = ea.data();
        d.off_a = oa.data();
        d.elem_b = eb.data();
        d.off_b = ob.data();
        d.is_used = 1;
        rc = exo_problem_add_task_map(p->shadow, &d);
        if (rc != EXO_OK) return rc;
    }
    rc = exo_problem_finalize(p->shadow);
    if (rc != EXO_OK) return rc;
    if (p->n_generated > 0) rc = exo_problem_set_generated(p->shadow, p->T, p->n_generated, p->h_generated.data());
    return rc;
}

int exo_problem_hessian_batch(exo_problem* p, const double* d_x, const double* d_jac, int64_t n_evals, int32_t T, int32_t layout, double* d_hess,
                              void* stream)
{
    if (!p) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (n_evals < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_evals == 0) return EXO_OK;
    if (!d_x || !d_jac || !d_hess) return fail(EXO_ERR_INVALID, "NULL buffer");
    if ((int)p->maps.size() > COST_MAXMAPS) return fail(EXO_ERR_UNSUPPORTED, "more task maps than the Hessian kernel supports");
    int rc = ensure_shadow(p);
    if (rc != EXO_OK) return rc;
    const int64_t N = p->prog.hdr.n_ctrl, LJ = p->prog.hdr.length_jac;
    const int64_t FP = p->shadow->prog.hdr.length_phi, FJ = p->shadow->prog.hdr.length_jac;  // 12 F, 6 F
    if (p->shadow_capacity < n_evals)
    {
        if (p->d_shadow_phi) cudaFree(p->d_shadow_phi);
        if (p->d_shadow_jac) cudaFree(p->d_shadow_jac);
        p->d_shadow_phi = p->d_shadow_jac = nullptr;
        p->shadow_capacity = 0;
        CUDA_TRY(cudaMalloc((void**)&p->d_shadow_phi, std::max<size_t>(8, (size_t)n_evals * FP * sizeof(double))));
        CUDA_TRY(cudaMalloc((void**)&p->d_shadow_jac, std::max<size_t>(8, (size_t)n_evals * FJ * N * sizeof(double))));
        p->shadow_capacity = n_evals;
    }
    // frames and their full Jacobians in SoA (coalesced for the Hessian kernel whatever the caller's layout is)
    if (FP > 0)
    {
        rc = launch(p->shadow, d_x, n_evals, n_evals, T, layout, p->d_shadow_phi, p->d_shadow_jac, (cudaStream_t)stream, false);
        if (rc != EXO_OK) return rc;
    }
    KernelParams k;
    memset(&k, 0, sizeof(k));
    rc = fill_strides(p, layout, n_evals, k);
    if (rc != EXO_OK) return rc;
    HessParams h;
    memset(&h, 0, sizeof(h));
    h.frames_phi = p->d_shadow_phi;
    h.frames_jac = p->d_shadow_jac;
    h.map_jac = d_jac;
    h.hess = d_hess;
    h.n_evals = n_evals;
    h.n_maps = (int)p->maps.size();
    h.n_ctrl = (int)N;
    h.mj_se = k.j_se; h.mj_sr = k.j_sr; h.mj_sc = k.j_sc;
    // the shadow problem was launched with the caller's layout for X; its outputs follow the same layout family
    KernelParams ks;
    memset(&ks, 0, sizeof(ks));
    rc = fill_strides(p->shadow, layout, n_evals, ks);
    if (rc != EXO_OK) return rc;
    h.fp_se = ks.phi_se; h.fp_sk = ks.phi_sk; h.fj_se = ks.j_se; h.fj_sr = ks.j_sr; h.fj_sc = ks.j_sc;
    if (layout == EXO_LAYOUT_SOA)
    {
        h.h_se = 1; h.h_sr = N * N * n_evals; h.h_si = N * n_evals; h.h_sj = n_evals;  // H[LJ][N][N][E]
    }
    else
    {
        h.h_se = LJ * N * N; h.h_sr = N * N; h.h_si = N; h.h_sj = 1;  // H[E][LJ][N][N]
    }
    for (size_t i = 0; i < p->maps.size(); ++i)
    {
        const MapSpec& m = p->maps[i];
        HessMapP& hm = h.map[i];
        hm.kind = m.kind;
        hm.is_used = m.is_used ? 1 : 0;
        hm.n_frames = m.n_frames;
        hm.frame_start = p->shadow_frame_start[i];
        hm.start_jacobian = m.start_jacobian;
        hm.length_jacobian = m.length_jacobian;
        hm.rows_per_frame = m.n_frames > 0 ? m.length_jacobian / m.n_frames : 0;
        if (m.kind == EXO_MAP_FRAMES) return fail(EXO_ERR_UNSUPPORTED, "raw-frame maps have no task-map Hessian");
    }
    cudaError_t e = launch_hessian(h, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "Hessian kernel launch");
    p->launches++;
    p->last_kernel = "exo_hessian_kernel";
    return EXO_OK;
}

// ---- batched IK -----------------------------------------------------------------------------------------------------------------
int exo_problem_ik_solve_batch(exo_problem* p, double* d_q, int64_t n_evals, const exo_ik_params* prm, double* d_cost, int32_t* d_iterations,
                               int32_t* d_termination, void* stream_)
{
    if (!p || !prm) return fail(EXO_ERR_INVALID, "NULL argument");
    if (!p->finalized) return fail(EXO_ERR_STATE, "problem not finalized");
    if (!p->cost_set) return fail(EXO_ERR_STATE, "exo_problem_set_cost has not been called");
    if (p->cost.T != 1) return fail(EXO_ERR_INVALID, "This IKSolver can't solve time-indexed problems (ik_solver.cpp:38-41)");
    if (n_evals < 0) return fail(EXO_ERR_INVALID, "negative batch size");
    if (n_evals == 0) return EXO_OK;
    if (!d_q || !d_cost || !d_iterations || !d_termination) return fail(EXO_ERR_INVALID, "NULL buffer");
    if (prm->max_iterations < 1) return fail(EXO_ERR_INVALID, "max_iterations must be >= 1");
    const int N = p->prog.hdr.n_ctrl, LP = p->prog.hdr.length_phi, LJ = p->prog.hdr.length_jac;
    if (LJ > IK_BIG_MAXLJ || N > IK_MAXN) return fail(EXO_ERR_UNSUPPORTED, "batched IK supports up to 64 task-space rows and 64 joints");
    const bool big = LJ > IK_MAXLJ || N > CH_MAXN;  // whole-body problems: one warp per configuration, the system in shared memory (exo_ik_step_warp_kernel)
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t E = n_evals;
    if (p->ik_capacity < E)
    {
        if (p->d_ik) cudaFree(p->d_ik);
        if (p->d_ik_i) cudaFree(p->d_ik_i);
        p->d_ik = nullptr;
        p->d_ik_i = nullptr;
        p->ik_capacity = 0;
        CUDA_TRY(cudaMalloc((void**)&p->d_ik, (size_t)E * (LP + (size_t)LJ * N + 2 * N + 4) * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&p->d_ik_i, ((size_t)3 * E + 8) * sizeof(int32_t)));
        p->ik_capacity = E;
    }
    double* w = p->d_ik;
    double* d_phi = w; w += (size_t)E * LP;
    double* d_jac = w; w += (size_t)E * LJ * N;
    IkParams K;
    memset(&K, 0, sizeof(K));
    K.n_evals = E;
    K.n_ctrl = N;
    K.length_jac = LJ;
    K.tolerance = prm->tolerance;
    K.step_tolerance = prm->step_tolerance;
    K.gradient_tolerance = prm->gradient_tolerance;
    K.regmax = prm->maximum_regularization;
    K.jac = d_jac;
    K.q = d_q;
    K.q_trial = w; w += (size_t)E * N;
    K.qd = w; w += (size_t)E * N;
    K.err_prev = w; w += E;
    K.stop = w; w += E;
    K.lambda = w; w += E;
    K.steplen = w; w += E;
    K.warp_step = big ? 1 : 0;
    K.single_iteration = prm->max_iterations == 1 ? 1 : 0;
    K.max_step = prm->max_step;
    K.th_stepinc = prm->threshold_regularization_increase;
    K.th_stepdec = prm->threshold_regularization_decrease;
    K.regmin = prm->minimum_regularization;
    const bool max_step_mode = K.single_iteration && K.max_step != 0.0;  // one clamped step instead of the line search
    K.cost = d_cost;
    K.active = p->d_ik_i;
    K.ls_done = p->d_ik_i + E;
    K.n_active = p->d_ik_i + 2 * E;
    K.n_list = p->d_ik_i + 2 * E + 4;
    K.list = p->d_ik_i + 2 * E + 8;
    K.iterations = d_iterations;
    K.termination = d_termination;
    for (size_t i = 0; i < p->maps.size(); ++i)
        for (int r = 0; r < p->maps[i].length_jacobian; ++r) K.row_rho[p->maps[i].start_jacobian + r] = p->cost_rho_t0[i];
    for (int c = 0; c < N; ++c) K.w_inv[c] = p->cost_W.empty() ? 1.0 : 1.0 / p->cost_W[c];  // W_inv_ = prob_->W.inverse() (ik_solver.cpp:45)
    // cost parameters bound to the workspace (SoA)
    CostParams C = p->cost;
    C.x = d_q;
    C.phi = d_phi;
    C.n_traj = E;
    C.x_se = 1; C.x_sk = E; C.phi_se = 1; C.phi_sk = E;
    C.W = nullptr;  // end-pose problems have no transition cost
    CUDA_TRY(launch_ik_init(K, prm->regularization_rate, stream));
    int rc = EXO_OK;
    if (p->compiled_version != p->tree->version)
    {
        rc = compile_and_upload(p);
        if (rc != EXO_OK) return rc;
    }
    // fused line search where the fused cost kernel applies (serial chains, EffFrame / EffPosition / JointLimit costs)
    ChainGradParams G;
    memset(&G, 0, sizeof(G));
    const bool fused_ls = fused_costgrad_eligible(p, G);
    ChainLineSearchParams LS;
    memset(&LS, 0, sizeof(LS));
    if (fused_ls)
    {
        G.rho = C.rho; G.goal = C.goal; G.goal_rot = C.goal_rot; G.W = nullptr;
        G.ct = C.ct;
        G.T = 1; G.t_first = 0; G.T_goal = C.T_goal; G.n_maps = C.n_maps; G.length_phi = C.length_phi; G.n_rot = C.n_rot;
        LS.q = d_q; LS.q_trial = K.q_trial; LS.qd = K.qd; LS.err_prev = K.err_prev; LS.cost = d_cost; LS.active = K.active; LS.ls_done = K.ls_done;
        LS.steplen = K.steplen;
    }
    // Chain programs: the line search of every iteration is one fused kernel (exo_problem_set_kernel(1) takes the trial /
    // update / accept launches instead).  EXO_IK_ONE_LAUNCH=1 runs the WHOLE solve as one launch, every lane iterating on its
    // own -- no host round trips, but a lane's 100 iterations are one serial instruction stream at 12 warps/SM: measured
    // 24.7 ms against 16.5 ms for 65 536 LWR starts, so it is opt-in (useful where the host must not be involved, e.g. inside
    // a captured graph).
    const char* one_env = getenv("EXO_IK_ONE_LAUNCH");
    const bool one_launch = fused_ls && one_env && one_env[0] == '1';
    if (one_launch)
    {
        chain_bind(p->chain, d_q, d_phi, d_jac, E, 1, E, 1, E, 1, (int64_t)N * E, E);
        CUDA_TRY(chain_ik_solve_launch(p->chain, G, LS, K, C, prm->max_iterations, stream));
        p->launches++;
    }
    for (int it = 0; it < (one_launch ? 0 : prm->max_iterations); ++it)
    {
        rc = launch(p, d_q, E, E, 1, EXO_LAYOUT_SOA, d_phi, d_jac, stream, false);  // prob_->Update(q_)
        if (rc != EXO_OK) return rc;
        CUDA_TRY(launch_ik_step(K, C, stream));
        if (max_step_mode)
        {
            // exo_ik_step_kernel has clamped the step and taken it (ScaleToStepSize; q_ -= qd_)
        }
        else if (fused_ls)
        {
            // chain programs: the ten trial states are evaluated in registers by one kernel
            chain_bind(p->chain, K.q_trial, nullptr, nullptr, E, 1, E, 1, E, 1, (int64_t)N * E, E);
            CUDA_TRY(chain_ik_linesearch_launch(p->chain, G, LS, stream));
            p->launches++;
        }
        else
        for (int ai = 0; ai < 10; ++ai)
        {
            const double alpha = ai == 0 ? 1.0 : 0.1 - (9 - ai) * ((0.1 - 1.0) / 9.0);  // Eigen::VectorXd::LinSpaced(10, 1.0, 0.1), see ik_alpha()
            CUDA_TRY(launch_ik_trial(K, alpha, ai == 0, stream));
            rc = launch(p, K.q_trial, E, E, 1, EXO_LAYOUT_SOA, d_phi, nullptr, stream, false);  // prob_->Update(q_tmp): FK only
            if (rc != EXO_OK) return rc;
            CUDA_TRY(launch_ik_accept(K, C, alpha, stream));
        }
        const bool last = it + 1 == prm->max_iterations;
        // the staged line search is thirty launches per iteration: ask every iteration whether anyone is still active; the fused chain
        // kernels make an iteration so cheap that every fourth is enough
        const bool poll = last || !fused_ls || (it % 4 == 3);
        if (poll) CUDA_TRY(cudaMemsetAsync(K.n_active, 0, sizeof(int32_t), stream));
        CUDA_TRY(launch_ik_finish(K, last ? 1 : 0, stream));
        if (poll && !last)
        {
            int32_t n_active = 0;
            CUDA_TRY(cudaMemcpyAsync(&n_active, K.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (n_active == 0) break;
        }
    }
    // cost at the returned states
    rc = launch(p, d_q, E, E, 1, EXO_LAYOUT_SOA, d_phi, nullptr, stream, false);
    if (rc != EXO_OK) return rc;
    return exo_problem_cost_batch(p, d_q, d_phi, E, EXO_LAYOUT_SOA, d_cost, nullptr, stream_);
}

}  // extern "C"
