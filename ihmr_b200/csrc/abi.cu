// extern "C" surface of libihmr_b200.so — see include/ihmr_b200.h for the contract.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "kernels.cuh"

namespace ihmr {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

size_t opt_ws_bytes(int B);
int opt_stage(const ihmr_model* m, int B, int bs_norm, float* params, const ihmr_targets_t* tg,
              const ihmr_stage_t* stg, int save_mid_freq, int optimizer, void* ws, cudaStream_t st);
int opt_value_and_grad(const ihmr_model* m, int B, int bs_norm, const float* params, const ihmr_targets_t* tg,
                       const ihmr_stage_t* stg, float* losses6, float* grad, void* ws, cudaStream_t st);
int opt_final(const ihmr_model* m, int B, const float* params, const ihmr_targets_t* tg, float* right_verts,
              float* left_verts, float* joints_3d, float* collision_loss, float* collision_origin,
              float* j3d_loss_p, void* ws, cudaStream_t st);
int opt_profile_iteration(const ihmr_model* m, int B, int bs_norm, float* params, const ihmr_targets_t* tg,
                          const ihmr_stage_t* stg, float* ms, void* ws, cudaStream_t st);
int select_snapshots(int S, int B, const float* crit, const ihmr_stage_t* stg, int* index, cudaStream_t st);
int opt_criteria(const ihmr_model* m, int B, const float* params, const ihmr_targets_t* tg, float w_joints_2d, float w_joints_3d,
                 float* criteria, void* ws, cudaStream_t st);

template <typename T>
static int upload(T** dst, const std::vector<T>& host) {
    IHMR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), host.size() * sizeof(T)));
    IHMR_CUDA_OK(cudaMemcpy(*dst, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return IHMR_OK;
}

// D rows 135..144 (shape directions, k-major) and the joint regression of the shape space
static void build_shape_rows(const float* shapedirs, const float* Jreg, std::vector<float>& D,
                             std::vector<float>& Js) {
    for (int k = 0; k < NB; ++k)
        for (int c = 0; c < NC; ++c) D[(size_t)(NPF + k) * LDN + c] = shapedirs[(size_t)c * NB + k];
    for (int k = 0; k < NB; ++k)
        for (int j = 0; j < NJ; ++j)
            for (int c = 0; c < 3; ++c) {
                double acc = 0.0;
                for (int v = 0; v < NV; ++v) acc += (double)Jreg[j * NV + v] * shapedirs[((size_t)v * 3 + c) * NB + k];
                Js[k * 48 + j * 3 + c] = (float)acc;
            }
}

// shape directions per vertex as 8 planes of float4: entry e = c * 10 + k of vertex v sits at
// Sv[((e / 4) * NV + v) * 4 + e % 4] (entries 30, 31 are padding)
static std::vector<float> build_sv(const float* shapedirs) {
    std::vector<float> sv((size_t)NV * 32, 0.f);
    for (int v = 0; v < NV; ++v)
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < NB; ++k) {
                const int e = c * NB + k;
                sv[((size_t)(e / 4) * NV + v) * 4 + e % 4] = shapedirs[((size_t)v * 3 + c) * NB + k];
            }
    return sv;
}

// Static spatial clustering of the faces (rest pose): recursive median split along the longest
// axis of the centroids into ceil(NF/32) groups of <= 32 faces.  The posed mesh is articulated,
// so rest-pose neighbours stay neighbours and the per-frame cluster boxes stay tight.
static void split_faces(std::vector<int>& ids, int lo, int hi, int groups, const std::vector<float>& cen,
                        std::vector<std::vector<int>>& out) {
    if (groups <= 1) {
        out.emplace_back(ids.begin() + lo, ids.begin() + hi);
        return;
    }
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; ++i)
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], cen[ids[i] * 3 + a]); mx[a] = std::max(mx[a], cen[ids[i] * 3 + a]); }
    int ax = 0;
    for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    const int gl = groups / 2;
    const int mid = lo + (int)((long long)(hi - lo) * gl / groups);
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi,
                     [&](int p, int q) { return cen[p * 3 + ax] < cen[q * 3 + ax] || (cen[p * 3 + ax] == cen[q * 3 + ax] && p < q); });
    split_faces(ids, lo, mid, gl, cen, out);
    split_faces(ids, mid, hi, groups - gl, cen, out);
}

static std::vector<uint16_t> cluster_table(const float* v_template, const int32_t* faces) {
    constexpr int ncl = (NF + 31) / 32;
    std::vector<float> cen(NF * 3);
    for (int f = 0; f < NF; ++f)
        for (int a = 0; a < 3; ++a)
            cen[f * 3 + a] = (v_template[faces[f * 3] * 3 + a] + v_template[faces[f * 3 + 1] * 3 + a] + v_template[faces[f * 3 + 2] * 3 + a]) / 3.f;
    std::vector<int> ids(NF);
    for (int f = 0; f < NF; ++f) ids[f] = f;
    std::vector<std::vector<int>> groups;
    split_faces(ids, 0, NF, ncl, cen, groups);
    std::vector<uint16_t> tab((size_t)ncl * 32 * 4, 0);
    for (int c = 0; c < ncl; ++c)
        for (size_t i = 0; i < groups[c].size() && i < 32; ++i) {
            const int f = groups[c][i];
            uint16_t* t = &tab[((size_t)c * 32 + i) * 4];
            t[0] = (uint16_t)faces[f * 3]; t[1] = (uint16_t)faces[f * 3 + 1]; t[2] = (uint16_t)faces[f * 3 + 2]; t[3] = 1;
        }
    return tab;
}

static std::vector<float> transpose_D(const std::vector<float>& D) {
    std::vector<float> DT((size_t)LDN * KP, 0.f);
    for (int k = 0; k < KP; ++k)
        for (int c = 0; c < LDN; ++c) DT[(size_t)c * KP + k] = D[(size_t)k * LDN + c];
    return DT;
}

}  // namespace ihmr

using namespace ihmr;

#define IHMR_CHECK_ARG(cond)                                              \
    do {                                                                  \
        if (!(cond)) {                                                    \
            ihmr::set_error("invalid argument: %s (%s)", #cond, __func__); \
            return IHMR_E_INVALID;                                        \
        }                                                                 \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

extern "C" {

const char* ihmr_last_error(void) { return g_err; }
int ihmr_abi_version(void) { return 1; }
unsigned long long ihmr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ihmr_model_create(const float* v_template, const float* shapedirs, const float* posedirs,
                      const float* J_regressor, const float* lbs_weights, const int32_t* parents,
                      const float* hands_mean, const int32_t* faces_right, const int32_t* faces_left,
                      int device, ihmr_model_t** out) {
    IHMR_CHECK_ARG(v_template && shapedirs && posedirs && J_regressor && lbs_weights && parents && hands_mean);
    IHMR_CHECK_ARG(faces_right && faces_left && out);
    for (int j = 0; j < NJ; ++j) IHMR_CHECK_ARG(j == 0 ? parents[j] < 0 : (parents[j] >= 0 && parents[j] < j));
    for (int i = 0; i < NF * 3; ++i) IHMR_CHECK_ARG(faces_right[i] >= 0 && faces_right[i] < NV && faces_left[i] >= 0 && faces_left[i] < NV);
    cudaDeviceProp prop;
    IHMR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", device, prop.major, prop.minor);
        return IHMR_E_ARCH;
    }
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cudaSetDevice(%d) failed", device); return IHMR_E_CUDA; }

    ihmr_model* m = new ihmr_model();
    memset(m, 0, sizeof(*m));
    m->sdf_box_scale = 0.6f;          // (1 + 0.2) / 2, +x parity ray: ihmr_model_set_sdf_conventions changes them
    m->sdf_ray_axis = 0;
    m->device = device;
    m->num_sms = prop.multiProcessorCount;
    for (int j = 0; j < NJ; ++j) m->parents[j] = parents[j];

    std::vector<float> D((size_t)KP * LDN, 0.f), vt(LDN, 0.f), Jt(48), Js(NB * 48), Wt(NJ * NV), W4(4 * NV * 4), hm(48, 0.f);
    for (int k = 0; k < NPF; ++k) memcpy(&D[(size_t)k * LDN], posedirs + (size_t)k * NC, NC * sizeof(float));
    build_shape_rows(shapedirs, J_regressor, D, Js);
    memcpy(vt.data(), v_template, NC * sizeof(float));
    for (int j = 0; j < NJ; ++j)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int v = 0; v < NV; ++v) acc += (double)J_regressor[j * NV + v] * v_template[v * 3 + c];
            Jt[j * 3 + c] = (float)acc;
        }
    for (int v = 0; v < NV; ++v)
        for (int j = 0; j < NJ; ++j) {
            Wt[j * NV + v] = lbs_weights[v * NJ + j];
            W4[((size_t)(j / 4) * NV + v) * 4 + (j % 4)] = lbs_weights[v * NJ + j];
        }
    memcpy(hm.data() + 3, hands_mean, 45 * sizeof(float));
    std::vector<float> DT = transpose_D(D);
    std::vector<float> DTq(gemm_presplit_floats(LDN, KP)), Dq(gemm_presplit_floats(KP, LDN));
    gemm_presplit_b(DT.data(), LDN, KP, KP, DTq.data());
    gemm_presplit_b(D.data(), KP, LDN, LDN, Dq.data());
    std::vector<float> Jreg(J_regressor, J_regressor + NJ * NV);
    std::vector<uint16_t> fr(NF * 4, 0), fl(NF * 4, 0);
    for (int f = 0; f < NF; ++f)
        for (int c = 0; c < 3; ++c) {
            fr[f * 4 + c] = (uint16_t)faces_right[f * 3 + c];
            fl[f * 4 + c] = (uint16_t)faces_left[f * 3 + c];
        }
    int rc;
    if ((rc = upload(&m->D, D)) || (rc = upload(&m->DT, DT)) || (rc = upload(&m->DTq, DTq)) || (rc = upload(&m->Dq, Dq)) ||
        (rc = upload(&m->vtemp, vt)) ||
        (rc = upload(&m->Jt, Jt)) || (rc = upload(&m->Js, Js)) || (rc = upload(&m->Wt, Wt)) ||
        (rc = upload(&m->W4, W4)) || (rc = upload(&m->hands_mean, hm)) || (rc = upload(&m->Jreg, Jreg)) ||
        (rc = upload(&m->Sv, build_sv(shapedirs))) ||
        (rc = upload(&m->faces[0], fr)) || (rc = upload(&m->faces[1], fl)) ||
        (rc = upload(&m->cl_tri[0], cluster_table(v_template, faces_right))) ||
        (rc = upload(&m->cl_tri[1], cluster_table(v_template, faces_left)))) {
        ihmr_model_destroy(m);
        return rc;
    }
    *out = m;
    return IHMR_OK;
}

void ihmr_model_destroy(ihmr_model_t* m) {
    if (!m) return;
    DeviceGuard guard(m->device);
    cudaFree(m->D); cudaFree(m->DT); cudaFree(m->DTq); cudaFree(m->Dq); cudaFree(m->vtemp); cudaFree(m->Jt); cudaFree(m->Js);
    cudaFree(m->Wt); cudaFree(m->W4); cudaFree(m->hands_mean); cudaFree(m->Jreg); cudaFree(m->Sv);
    cudaFree(m->faces[0]); cudaFree(m->faces[1]); cudaFree(m->cl_tri[0]); cudaFree(m->cl_tri[1]);
    delete m;
}

int ihmr_model_set_sdf_conventions(ihmr_model_t* m, float scale_factor, int ray_axis) {
    IHMR_CHECK_ARG(m && scale_factor > -1.0f && scale_factor < 10.0f && ray_axis >= 0 && ray_axis <= 2);
    m->sdf_box_scale = (1.0f + scale_factor) * 0.5f;
    m->sdf_ray_axis = ray_axis;
    return IHMR_OK;
}

int ihmr_model_update_shapedirs(ihmr_model_t* m, const float* shapedirs, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && shapedirs);
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<float> Jreg(NJ * NV), D((size_t)KP * LDN, 0.f), Js(NB * 48);
    IHMR_CUDA_OK(cudaMemcpyAsync(Jreg.data(), m->Jreg, Jreg.size() * 4, cudaMemcpyDeviceToHost, st));
    IHMR_CUDA_OK(cudaMemcpyAsync(D.data(), m->D, D.size() * 4, cudaMemcpyDeviceToHost, st));
    IHMR_CUDA_OK(cudaStreamSynchronize(st));   // model mutation is the one synchronising call
    build_shape_rows(shapedirs, Jreg.data(), D, Js);
    std::vector<float> DT = transpose_D(D);
    std::vector<float> DTq(gemm_presplit_floats(LDN, KP)), Dq(gemm_presplit_floats(KP, LDN));
    gemm_presplit_b(DT.data(), LDN, KP, KP, DTq.data());
    gemm_presplit_b(D.data(), KP, LDN, LDN, Dq.data());
    IHMR_CUDA_OK(cudaMemcpyAsync(m->D, D.data(), D.size() * 4, cudaMemcpyHostToDevice, st));
    IHMR_CUDA_OK(cudaMemcpyAsync(m->DT, DT.data(), DT.size() * 4, cudaMemcpyHostToDevice, st));
    IHMR_CUDA_OK(cudaMemcpyAsync(m->DTq, DTq.data(), DTq.size() * 4, cudaMemcpyHostToDevice, st));
    IHMR_CUDA_OK(cudaMemcpyAsync(m->Dq, Dq.data(), Dq.size() * 4, cudaMemcpyHostToDevice, st));
    IHMR_CUDA_OK(cudaMemcpyAsync(m->Js, Js.data(), Js.size() * 4, cudaMemcpyHostToDevice, st));
    const std::vector<float> sv = build_sv(shapedirs);
    IHMR_CUDA_OK(cudaMemcpyAsync(m->Sv, sv.data(), sv.size() * 4, cudaMemcpyHostToDevice, st));
    IHMR_CUDA_OK(cudaStreamSynchronize(st));
    return IHMR_OK;
}

int ihmr_gemm_tf32x3(int M, int Nc, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     ihmr_stream_t stream) {
    IHMR_CHECK_ARG(A && B && C && M >= 0);
    return launch_gemm_tf32x3(M, Nc, K, A, lda, B, ldb, C, ldc, static_cast<cudaStream_t>(stream));
}

int ihmr_gemm_reference_fp32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                             ihmr_stream_t stream) {
    IHMR_CHECK_ARG(A && B && C && M >= 0 && K % 8 == 0 && N % 4 == 0);
    return launch_sgemm_reference(M, N, K, A, lda, B, ldb, C, ldc, static_cast<cudaStream_t>(stream));
}

size_t ihmr_mano_workspace_bytes(int n_hands) { return n_hands > 0 ? mano_ws_bytes(n_hands) : 0; }

int ihmr_mano_forward(const ihmr_model_t* m, int n, const float* global_orient, const float* hand_pose,
                      const float* betas, float* vertices, float* joints, void* workspace,
                      size_t workspace_bytes, ihmr_stream_t stream) {
    NvtxRange range("ihmr_mano_forward");
    IHMR_CHECK_ARG(m && n >= 0 && global_orient && hand_pose && betas && vertices && workspace);
    if (workspace_bytes < mano_ws_bytes(n)) { set_error("workspace too small: %zu < %zu", workspace_bytes, mano_ws_bytes(n)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ManoWs w = mano_ws_carve(workspace, n);
    HandSrc src;
    src.orient = global_orient; src.pose = hand_pose; src.betas = betas;
    int rc;
    if ((rc = launch_pose_prep(m, n, src, w.X, w.A, joints ? joints : w.joints, st))) return rc;
    if ((rc = launch_blend_fwd(m, n, w.X, w.off, st))) return rc;
    return launch_skin_fwd(m, n, w.off, w.A, vertices, st);
}

int ihmr_mano_backward(const ihmr_model_t* m, int n, const float* global_orient, const float* hand_pose,
                       const float* betas, const float* grad_vertices, const float* grad_joints,
                       float* grad_global_orient, float* grad_hand_pose, float* grad_betas,
                       void* workspace, size_t workspace_bytes, ihmr_stream_t stream) {
    NvtxRange range("ihmr_mano_backward");
    IHMR_CHECK_ARG(m && n >= 0 && global_orient && hand_pose && betas && workspace);
    IHMR_CHECK_ARG(grad_global_orient && grad_hand_pose && grad_betas);
    if (workspace_bytes < mano_ws_bytes(n)) { set_error("workspace too small: %zu < %zu", workspace_bytes, mano_ws_bytes(n)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ManoWs w = mano_ws_carve(workspace, n);
    HandSrc src;
    src.orient = global_orient; src.pose = hand_pose; src.betas = betas;
    int rc;
    if ((rc = launch_pose_prep(m, n, src, w.X, w.A, w.joints, st))) return rc;
    if ((rc = launch_blend_fwd(m, n, w.X, w.off, st))) return rc;
    if ((rc = launch_skin_bwd(m, n, w.off, w.A, grad_vertices, nullptr, w.gposed, w.dA, st))) return rc;
    if ((rc = launch_blend_bwd(m, n, w.gposed, w.dX, st, SparseGrad(), w.off))) return rc;
    HandGrad hg;
    hg.orient = grad_global_orient; hg.pose = grad_hand_pose; hg.betas = grad_betas;
    return launch_pose_bwd(m, n, src, w.dA, grad_joints, w.dX, hg, st);
}

size_t ihmr_sdf_workspace_bytes(int n_frames) { return n_frames > 0 ? sdf_ws_bytes(n_frames) : 0; }

int ihmr_sdf_loss(const ihmr_model_t* m, int n_frames, const float* hand_verts, float* losses, float* per_vert,
                  float* origin_scale, float* grad_hand_verts, float robustifier, void* workspace,
                  size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && n_frames >= 0 && hand_verts && losses && (workspace || n_frames == 0));
    if (workspace_bytes < ihmr_sdf_workspace_bytes(n_frames)) { set_error("workspace too small: %zu < %zu", workspace_bytes, ihmr_sdf_workspace_bytes(n_frames)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    SdfArgs a;
    a.verts = hand_verts; a.losses = losses; a.per_vert = per_vert; a.origin = origin_scale;
    a.gverts = grad_hand_verts; a.robustifier = robustifier; a.ws = workspace;
    return launch_sdf(m, n_frames, a, static_cast<cudaStream_t>(stream));
}

int ihmr_sdf_loss_exact(const ihmr_model_t* m, int n_frames, const float* hand_verts, float* losses, float* per_vert,
                        float* origin_scale, float* grad_hand_verts, void* workspace, size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && n_frames >= 0 && hand_verts && losses && (workspace || n_frames == 0));
    if (workspace_bytes < ihmr_sdf_workspace_bytes(n_frames)) { set_error("workspace too small: %zu < %zu", workspace_bytes, ihmr_sdf_workspace_bytes(n_frames)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    SdfArgs a;
    a.verts = hand_verts; a.losses = losses; a.per_vert = per_vert; a.origin = origin_scale;
    a.gverts = grad_hand_verts; a.ws = workspace;
    return launch_sdf_exact(m, n_frames, a, static_cast<cudaStream_t>(stream));
}

int ihmr_sdf_stats(const ihmr_model_t* m, int n_frames, const float* hand_verts, float* losses, int* stats,
                   void* workspace, size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && n_frames >= 0 && hand_verts && losses && stats && (workspace || n_frames == 0));
    if (workspace_bytes < ihmr_sdf_workspace_bytes(n_frames)) { set_error("workspace too small: %zu < %zu", workspace_bytes, ihmr_sdf_workspace_bytes(n_frames)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    SdfArgs a;
    a.verts = hand_verts; a.losses = losses; a.stats = stats; a.ws = workspace;
    return launch_sdf(m, n_frames, a, static_cast<cudaStream_t>(stream));
}

int ihmr_eval_metrics(int n_frames, const float* pred_joints_3d, const float* gt_joints_3d, const float* collision_origin_scale,
                      const float* scale, float* out, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(n_frames >= 0 && pred_joints_3d && gt_joints_3d && collision_origin_scale && out);
    return launch_eval_metrics(n_frames, pred_joints_3d, gt_joints_3d, collision_origin_scale, scale, out, static_cast<cudaStream_t>(stream));
}

static int check_targets(const ihmr_targets_t* t) {
    IHMR_CHECK_ARG(t && t->init_joints_2d && t->init_joints_3d && t->init_hand_trans_j && t->gt_joints_3d && t->hand_type_array);
    return IHMR_OK;
}

static int check_stage(const ihmr_stage_t* s) {
    IHMR_CHECK_ARG(s && s->epoch >= 0 && s->n_filters >= 0 && s->n_filters <= 4);
    IHMR_CHECK_ARG(s->select_loss >= 0 && s->select_loss <= 2);
    for (int f = 0; f < s->n_filters; ++f) IHMR_CHECK_ARG(s->filter_loss[f] >= 0 && s->filter_loss[f] <= 2);
    return IHMR_OK;
}

int ihmr_mlp_input(int n, const float* img_feat, const float* params, float* x, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(n >= 0 && img_feat && params && x);
    return launch_mlp_input(n, img_feat, params, x, static_cast<cudaStream_t>(stream));
}

int ihmr_linear(int n, int in_dim, int out_dim, const float* x, int ldx, const float* weight, const float* bias, int relu,
                float* y, int ldy, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(n >= 0 && x && weight && y && in_dim > 0 && out_dim > 0 && in_dim % 32 == 0 && ldx % 4 == 0 && ldy % 4 == 0);
    IHMR_CHECK_ARG(ldx >= in_dim && ldy >= ((out_dim + 3) & ~3));
    return launch_linear(n, in_dim, out_dim, x, ldx, weight, bias, relu, y, ldy, static_cast<cudaStream_t>(stream));
}

int ihmr_mlp_apply(int n, const float* residual, int ldr, int n_segments, const int32_t* seg_col, const int32_t* seg_len,
                   const float* params_in, float* params_out, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(n >= 0 && residual && params_in && params_out && n_segments >= 0 && n_segments <= 8 && (n_segments == 0 || (seg_col && seg_len)));
    int tot = 0;
    for (int i = 0; i < n_segments; ++i) { IHMR_CHECK_ARG(seg_col[i] >= 0 && seg_len[i] > 0 && seg_col[i] + seg_len[i] <= IHMR_PARAM_DIM); tot += seg_len[i]; }
    IHMR_CHECK_ARG(tot <= ldr);
    return launch_mlp_apply(n, residual, ldr, n_segments, seg_col, seg_len, params_in, params_out, static_cast<cudaStream_t>(stream));
}

int ihmr_select_better(int n, const float* cur_criteria, float* prev_criteria, const ihmr_stage_t* stage, const float* new_params,
                       float* params, int32_t* kept, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(n >= 0 && cur_criteria && prev_criteria && new_params && params);
    int rc;
    if ((rc = check_stage(stage))) return rc;
    return launch_select_better(n, cur_criteria, prev_criteria, stage, new_params, params, kept, static_cast<cudaStream_t>(stream));
}

int ihmr_measure_fp32_peak(const ihmr_model_t* m, float* tflops, void* scratch, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && tflops && scratch);
    DeviceGuard guard(m->device);
    return measure_fp32_peak(m->num_sms, tflops, static_cast<float*>(scratch), static_cast<cudaStream_t>(stream));
}

size_t ihmr_opt_workspace_bytes(int n_frames) { return n_frames > 0 ? opt_ws_bytes(n_frames) : 0; }

int ihmr_opt_stage(const ihmr_model_t* m, int B, int bs_norm, float* params, const ihmr_targets_t* targets,
                   const ihmr_stage_t* stage, int save_mid_freq, int optimizer, void* workspace,
                   size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && B > 0 && bs_norm > 0 && params && workspace && save_mid_freq > 0);
    IHMR_CHECK_ARG(optimizer == IHMR_OPT_ADAM || optimizer == IHMR_OPT_SGD);
    int rc;
    if ((rc = check_targets(targets)) || (rc = check_stage(stage))) return rc;
    if (workspace_bytes < opt_ws_bytes(B)) { set_error("workspace too small: %zu < %zu", workspace_bytes, opt_ws_bytes(B)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    return opt_stage(m, B, bs_norm, params, targets, stage, save_mid_freq, optimizer, workspace, static_cast<cudaStream_t>(stream));
}

int ihmr_opt_final(const ihmr_model_t* m, int B, const float* params, const ihmr_targets_t* targets,
                   float* right_verts, float* left_verts, float* joints_3d, float* collision_loss,
                   float* collision_origin_scale, float* joints_3d_loss_p, void* workspace,
                   size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && B > 0 && params && workspace);
    int rc;
    if ((rc = check_targets(targets))) return rc;
    if (workspace_bytes < opt_ws_bytes(B)) { set_error("workspace too small: %zu < %zu", workspace_bytes, opt_ws_bytes(B)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    return opt_final(m, B, params, targets, right_verts, left_verts, joints_3d, collision_loss,
                     collision_origin_scale, joints_3d_loss_p, workspace, static_cast<cudaStream_t>(stream));
}

int ihmr_opt_criteria(const ihmr_model_t* m, int B, const float* params, const ihmr_targets_t* targets, float w_joints_2d,
                      float w_joints_3d, float* criteria, void* workspace, size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && B > 0 && params && workspace && criteria);
    int rc;
    if ((rc = check_targets(targets))) return rc;
    if (workspace_bytes < opt_ws_bytes(B)) { set_error("workspace too small: %zu < %zu", workspace_bytes, opt_ws_bytes(B)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    return opt_criteria(m, B, params, targets, w_joints_2d, w_joints_3d, criteria, workspace, static_cast<cudaStream_t>(stream));
}

int ihmr_opt_value_and_grad(const ihmr_model_t* m, int B, int bs_norm, const float* params,
                            const ihmr_targets_t* targets, const ihmr_stage_t* stage, float* losses6,
                            float* grad, void* workspace, size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && B > 0 && bs_norm > 0 && params && workspace);
    int rc;
    if ((rc = check_targets(targets)) || (rc = check_stage(stage))) return rc;
    if (workspace_bytes < opt_ws_bytes(B)) { set_error("workspace too small: %zu < %zu", workspace_bytes, opt_ws_bytes(B)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    return opt_value_and_grad(m, B, bs_norm, params, targets, stage, losses6, grad, workspace, static_cast<cudaStream_t>(stream));
}

int ihmr_select_snapshots(int n_snapshots, int n_frames, const float* criteria, const ihmr_stage_t* stage, int32_t* index,
                          ihmr_stream_t stream) {
    IHMR_CHECK_ARG(n_snapshots > 0 && n_frames >= 0 && criteria && index);
    int rc;
    if ((rc = check_stage(stage))) return rc;
    if (n_frames == 0) return IHMR_OK;
    return select_snapshots(n_snapshots, n_frames, criteria, stage, index, static_cast<cudaStream_t>(stream));
}

int ihmr_opt_profile_iteration(const ihmr_model_t* m, int B, int bs_norm, float* params,
                               const ihmr_targets_t* targets, const ihmr_stage_t* stage, float* ms_per_kernel,
                               void* workspace, size_t workspace_bytes, ihmr_stream_t stream) {
    IHMR_CHECK_ARG(m && B > 0 && bs_norm > 0 && params && workspace && ms_per_kernel);
    int rc;
    if ((rc = check_targets(targets)) || (rc = check_stage(stage))) return rc;
    if (workspace_bytes < opt_ws_bytes(B)) { set_error("workspace too small: %zu < %zu", workspace_bytes, opt_ws_bytes(B)); return IHMR_E_WORKSPACE; }
    DeviceGuard guard(m->device);
    return opt_profile_iteration(m, B, bs_norm, params, targets, stage, ms_per_kernel, workspace, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
