// multi.cu — what sits between the C ABI and the engines of api.cu:
//   * lanes: extra engines (own stream, own scratch, own worker thread) on the context's device.  A synchronous fused
//     call splits its scene batch into sub-batches that run concurrently on the lanes, so that one sub-batch's
//     latency-bound refit/LM kernel runs under another's match / solve / score kernels; cvg_detect_scenes_submit
//     hands whole batches to the lanes, so that a single-threaded caller pipelines calls (submit k+1, wait k).
//   * multi-device contexts (cvg_create_multi): one device context per listed GPU of the host, a resident model replica
//     per device, scene batches dealt to the devices by cost (pair sharding, SURVEY 8e-1: no data-path collective), and
//     cvg_match_knn2_sharded: train-tile sharding with ONE exchange step (ncclAllGather of 16 B per query and device
//     over NVLink) followed by the lexicographic (distance, index) merge.
// Reference seam: the loop nest of detectObjects (src/TestsDetector.cpp:38,58,99) under processAllTestImages
// (src/Output.cpp:23-57), which carries no state from one (view, scene) pair to the next.
#include "ctx.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

using namespace cvg;

// ---- NCCL, resolved at run time (a process that already loaded NCCL, e.g. through torch, shares that copy) --------
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : { "libnccl.so.2", "libnccl.so" }) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return;
        api.CommInitAll = (int (*)(ncclComm_t*, int, const int*))dlsym(h, "ncclCommInitAll");
        api.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
        api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
        api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd;
    });
    return api;
}
constexpr int NCCL_CHAR = 0;                               // ncclInt8 / ncclChar

struct MultiState {
    std::vector<int> devices;
    std::vector<cvg_ctx*> subs;                            // one device context (engine + lanes) per listed device
    std::vector<Worker*> workers;                          // one host thread per device for blocking per-device work
    std::vector<ncclComm_t> comms;                         // empty: exchange by device-to-device copies
    std::vector<DevBuf> loc_d, loc_i, all_d, all_i, qbuf, tbuf, out_idx, out_dist, out_acc;
    std::vector<cudaEvent_t> ev;
    unsigned next_dev = 0;
    const char* exchange = "none";
};

struct ChunkTask { int s0 = 0, s1 = 0; Done done; std::vector<float> pool; std::vector<int32_t> cnt; int match_path = 0; };

struct cvg_job {
    // single-device part
    cvg_ctx* ctx = nullptr; const cvg_models* m = nullptr; cvg_scenes* sc = nullptr;
    std::vector<ChunkTask*> chunks;
    std::vector<float> scales; bool have_scales = false;
    cvg_detect_params params;
    cvg_pair_result* per_pair = nullptr; float* inl_xy = nullptr; int64_t* inl_off = nullptr;
    // multi-device part: one sub-job per involved device, results scattered back into the caller's order at wait
    std::vector<cvg_job*> subs; std::vector<int> sub_dev;
    std::vector<std::vector<cvg_pair_result>> sub_res; std::vector<std::vector<float>> sub_xy; std::vector<std::vector<int64_t>> sub_off;
    const cvg_scenes* shell = nullptr; const cvg_models* shell_m = nullptr;
    ~cvg_job() { for (ChunkTask* t : chunks) delete t; for (cvg_job* j : subs) delete j; }
};

cvg_ctx* cvg_primary(cvg_ctx* c) { return (c && c->multi) ? c->multi->subs[0] : c; }
const cvg_ctx* cvg_primary(const cvg_ctx* c) { return (c && c->multi) ? c->multi->subs[0] : c; }
const cvg_models* cvg_models_on(const cvg_models* m, const cvg_ctx* eng)
{
    if (!m || m->replicas.empty() || !eng) return m;
    for (const cvg_models* r : m->replicas) if (r && r->device == eng->device) return r;
    return m->replicas[0];
}

// ---- lanes ------------------------------------------------------------------------------------------------
static int default_lanes()
{
    static const int n = [] { const char* e = getenv("CVG_LANES"); const int v = e ? atoi(e) : 3; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    return n;
}
static int lanes_of(const cvg_ctx* c) { return c->lanes_cfg > 0 ? c->lanes_cfg : default_lanes(); }

static int ensure_lanes(cvg_ctx* c, int n)
{
    while ((int)c->lanes.size() < n) {
        cvg_ctx* e = nullptr;
        const int rc = eng_create(&e, c->device, c->flags);
        if (rc) return rc;
        e->parent = c;
        Lane* l = new Lane();
        l->eng = e; l->worker = new Worker();
        c->lanes.push_back(l);
    }
    return CVG_OK;
}

static void destroy_lanes(cvg_ctx* c)
{
    for (Lane* l : c->lanes) { delete l->worker; eng_destroy(l->eng); delete l; }    // the worker drains its queue first
    c->lanes.clear();
}

// Sub-batches of a synchronous call: consecutive scene ranges of about equal cost (train rows), at most one per lane,
// and none at all for small calls (a split costs a thread hop per lane and shortens every launch).
static void plan_chunks(const cvg_ctx* c, const cvg_models* m, const cvg_scenes* sc, bool split, std::vector<std::pair<int, int>>& out)
{
    const int S = sc->ts.n_segs;
    out.clear();
    // a synchronous call gains nothing beyond three sub-batches (each ends in its own refit/LM chain; measured: 8 sub-batches
    // of 8 scenes halve the rate), pipelined jobs use every lane
    static const int max_split = [] { const char* e = getenv("CVG_SPLIT_LANES"); const int v = e ? atoi(e) : 3; return v < 1 ? 1 : v; }();
    int K = split ? std::min(std::min(lanes_of(c), max_split), S) : 1;
    const int64_t min_cost = c->split_min_cost >= 0 ? c->split_min_cost
                                                    : (getenv("CVG_SPLIT_MIN_COST") ? atoll(getenv("CVG_SPLIT_MIN_COST")) : (int64_t)1 << 25);
    const int64_t cost = sc->ts.rows_total * (int64_t)std::max(m->n_rows, 1);
    if (K > 1 && cost / K < min_cost) K = (int)std::max<int64_t>(1, cost / std::max<int64_t>(min_cost, 1));
    if (K <= 1) { out.push_back({ 0, S }); return; }
    int s = 0; int64_t done = 0;
    for (int k = 0; k < K && s < S; k++) {
        const int64_t target = sc->ts.rows_total * (k + 1) / K;
        int e = s;
        while (e < S && (done < target || e == s) && S - e > K - 1 - k) { done += sc->ts.segs[(size_t)e].rows; e++; }
        if (k == K - 1) e = S;
        if (e > s) out.push_back({ s, e });
        s = e;
    }
    if (s < S) out.back().second = S;
}

// Enqueue scenes x views of one device context on its lanes; returns without waiting.
static int dev_submit(cvg_ctx* c, const cvg_models* m, cvg_scenes* sc, const float* scales, const cvg_detect_params* p,
                      cvg_pair_result* per_pair, float* inl_xy, int64_t* inl_off, bool split, cvg_job** out)
{
    *out = nullptr;
    cvg_job* j = new cvg_job();
    j->ctx = c; j->m = m; j->sc = sc; j->params = *p; j->per_pair = per_pair; j->inl_xy = inl_xy; j->inl_off = inl_off;
    const int S = sc->ts.n_segs, V = m->n_views;
    if (scales) { j->scales.assign(scales, scales + S); j->have_scales = true; }
    std::vector<std::pair<int, int>> ranges;
    plan_chunks(c, m, sc, split, ranges);
    const int K = (int)ranges.size();
    const int need = split ? K : lanes_of(c);
    int rc = ensure_lanes(c, need);
    if (rc) { delete j; return rc; }
    const bool want_inl = inl_xy != nullptr && inl_off != nullptr;
    for (int k = 0; k < K; k++) {
        ChunkTask* t = new ChunkTask();
        t->s0 = ranges[(size_t)k].first; t->s1 = ranges[(size_t)k].second;
        j->chunks.push_back(t);
        // a split call occupies lanes 0..K-1 at once; pipelined jobs take the lanes in turn
        Lane* lane = c->lanes[split ? (size_t)k : (size_t)(c->next_lane++ % (unsigned)need)];
        cvg_ctx* eng = lane->eng;
        const int share = split ? K : need;
        lane->worker->post([=] {
            eng->wave_div = share;
            eng->blocking_sync = !split && need > 4;     // many lanes: their workers sleep while the GPU works (api.cu, sync_and_check)
            const int r = eng_detect_range(eng, m, sc, t->s0, t->s1, j->have_scales ? j->scales.data() : nullptr, &j->params,
                                           j->per_pair + (size_t)t->s0 * V, want_inl ? &t->pool : nullptr, want_inl ? &t->cnt : nullptr);
            t->match_path = eng->last_match_path;
            t->done.set(r, cvg_last_error());
        });
    }
    *out = j;
    return CVG_OK;
}

static int dev_wait(cvg_job* j)
{
    int rc = CVG_OK; std::string err;
    for (ChunkTask* t : j->chunks) {
        const int r = t->done.wait();
        if (r && !rc) { rc = r; err = t->done.err; }
    }
    if (rc) return cvg_set_err(rc, "%s", err.c_str());
    for (ChunkTask* t : j->chunks) j->ctx->last_match_path = std::max(t == j->chunks[0] ? 0 : j->ctx->last_match_path, t->match_path);
    if (j->inl_xy && j->inl_off) {
        // every chunk comes back packed pair after pair (eng_detect_range): concatenate in scene order
        const int V = j->m->n_views;
        int64_t o = 0;
        for (ChunkTask* t : j->chunks) {
            memcpy(j->inl_xy + 2 * (size_t)o, t->pool.data(), t->pool.size() * 4);
            for (size_t lp = 0; lp < t->cnt.size(); lp++) { j->inl_off[(size_t)t->s0 * V + lp] = o; o += t->cnt[lp]; }
        }
        j->inl_off[(size_t)j->sc->ts.n_segs * V] = o;
    }
    return CVG_OK;
}

// ---- multi-device: submit / wait ----------------------------------------------------------------------------
static int multi_submit(cvg_ctx* c, const cvg_models* m, const cvg_scenes* sc, const float* scales, const cvg_detect_params* p,
                        cvg_pair_result* per_pair, float* inl_xy, int64_t* inl_off, bool split, cvg_job** out)
{
    MultiState* ms = c->multi;
    *out = nullptr;
    if (m->replicas.size() != ms->subs.size() || sc->subs.size() != ms->subs.size())
        return cvg_set_err(CVG_ERR_INVALID, "models / scenes were not uploaded through this multi-device context");
    cvg_job* j = new cvg_job();
    j->ctx = c; j->shell = sc; j->shell_m = m; j->per_pair = per_pair; j->inl_xy = inl_xy; j->inl_off = inl_off;
    const int V = m->n_views;
    const bool want_inl = inl_xy != nullptr && inl_off != nullptr;
    const size_t nd = ms->subs.size();
    j->sub_res.resize(nd); j->sub_xy.resize(nd); j->sub_off.resize(nd);
    for (size_t d = 0; d < nd; d++) {
        cvg_scenes* sub = sc->subs[d];
        if (!sub) continue;
        const int Sd = sub->ts.n_segs;
        std::vector<float> sc_d;
        if (scales) { sc_d.resize((size_t)Sd); for (int s = 0; s < sc->n_scenes_total; s++) if (sc->dev_of_scene[(size_t)s] == (int)d) sc_d[(size_t)sc->local_of_scene[(size_t)s]] = scales[s]; }
        j->sub_res[d].resize((size_t)std::max(Sd * V, 1));
        if (want_inl) { j->sub_xy[d].resize((size_t)Sd * std::max(m->n_rows, 1) * 2); j->sub_off[d].resize((size_t)Sd * V + 1); }
        cvg_job* sj = nullptr;
        const int rc = dev_submit(ms->subs[d], m->replicas[d], sub, scales ? sc_d.data() : nullptr, p, j->sub_res[d].data(),
                                  want_inl ? j->sub_xy[d].data() : nullptr, want_inl ? j->sub_off[d].data() : nullptr, split, &sj);
        if (rc) { for (cvg_job* q : j->subs) dev_wait(q); delete j; return rc; }
        j->subs.push_back(sj); j->sub_dev.push_back((int)d);
    }
    *out = j;
    return CVG_OK;
}

static int multi_wait(cvg_job* j)
{
    int rc = CVG_OK; std::string err;
    for (cvg_job* sj : j->subs) {
        const int r = dev_wait(sj);
        if (r && !rc) { rc = r; err = cvg_last_error(); }
    }
    if (rc) return cvg_set_err(rc, "%s", err.c_str());
    const cvg_scenes* sc = j->shell; const cvg_models* m = j->shell_m;
    const int V = m->n_views, S = sc->n_scenes_total;
    const bool want_inl = j->inl_xy && j->inl_off;
    int64_t o = 0;
    for (int s = 0; s < S; s++) {
        const size_t d = (size_t)sc->dev_of_scene[(size_t)s]; const int ls = sc->local_of_scene[(size_t)s];
        memcpy(j->per_pair + (size_t)s * V, j->sub_res[d].data() + (size_t)ls * V, (size_t)V * sizeof(cvg_pair_result));
        if (!want_inl) continue;
        const int64_t a = j->sub_off[d][(size_t)ls * V], b = j->sub_off[d][(size_t)(ls + 1) * V];
        memcpy(j->inl_xy + 2 * (size_t)o, j->sub_xy[d].data() + 2 * (size_t)a, (size_t)(b - a) * 8);
        for (int v = 0; v < V; v++) j->inl_off[(size_t)s * V + v] = o + (j->sub_off[d][(size_t)ls * V + v] - a);
        o += b - a;
    }
    if (want_inl) j->inl_off[(size_t)S * V] = o;
    return CVG_OK;
}

static void multi_destroy(cvg_ctx* c)
{
    MultiState* ms = c->multi;
    if (!ms) return;
    for (Worker* w : ms->workers) delete w;
    NcclApi& na = nccl_api();
    for (size_t d = 0; d < ms->comms.size(); d++) if (ms->comms[d]) { cudaSetDevice(ms->devices[d]); na.CommDestroy(ms->comms[d]); }
    for (size_t d = 0; d < ms->subs.size(); d++) {
        cudaSetDevice(ms->devices[d]);
        for (std::vector<DevBuf>* v : { &ms->loc_d, &ms->loc_i, &ms->all_d, &ms->all_i, &ms->qbuf, &ms->tbuf, &ms->out_idx, &ms->out_dist, &ms->out_acc })
            if (d < v->size()) (*v)[d].release();
        if (d < ms->ev.size() && ms->ev[d]) cudaEventDestroy(ms->ev[d]);
        if (ms->subs[d]) { destroy_lanes(ms->subs[d]); eng_destroy(ms->subs[d]); }
    }
    delete ms;
    c->multi = nullptr;
}

extern "C" {

// ---- context ------------------------------------------------------------------------------------------------
int cvg_create(cvg_ctx** out, int device, unsigned flags) { return eng_create(out, device, flags); }

int cvg_create_multi(cvg_ctx** out, const int* devices, int n_devices, unsigned flags)
{
    if (!out || !devices || n_devices < 1 || n_devices > 64) return cvg_set_err(CVG_ERR_INVALID, "cvg_create_multi: bad argument");
    *out = nullptr;
    cvg_ctx* shell = new cvg_ctx();
    shell->device = -1; shell->flags = flags;
    MultiState* ms = new MultiState();
    shell->multi = ms;
    ms->devices.assign(devices, devices + n_devices);
    const size_t nd = (size_t)n_devices;
    ms->subs.assign(nd, nullptr); ms->ev.assign(nd, nullptr);
    for (std::vector<DevBuf>* v : { &ms->loc_d, &ms->loc_i, &ms->all_d, &ms->all_i, &ms->qbuf, &ms->tbuf, &ms->out_idx, &ms->out_dist, &ms->out_acc }) v->resize(nd);
    for (size_t d = 0; d < nd; d++) {
        int rc = eng_create(&ms->subs[d], devices[d], flags);
        if (!rc && cudaEventCreateWithFlags(&ms->ev[d], cudaEventDisableTiming) != cudaSuccess) rc = cvg_set_err(CVG_ERR_CUDA, "cudaEventCreate failed");
        if (rc) { multi_destroy(shell); delete shell; return rc; }
        ms->subs[d]->parent = shell;
        ms->workers.push_back(new Worker());
    }
    // One NCCL communicator per device for the exchange step of the train-tile sharded match.  NCCL refuses a
    // communicator that names a GPU twice (logical devices on one GPU, as the single-GPU tests use): the exchange then
    // runs as device-to-device copies, which is also the fallback when libnccl cannot be loaded.
    bool distinct = true;
    for (size_t a = 0; a < nd; a++) for (size_t b = a + 1; b < nd; b++) if (devices[a] == devices[b]) distinct = false;
    ms->exchange = "memcpy";
    const char* no_nccl = getenv("CVG_NO_NCCL");
    if (nd > 1 && distinct && !(no_nccl && atoi(no_nccl)) && nccl_api().ok) {
        ms->comms.assign(nd, nullptr);
        const int r = nccl_api().CommInitAll(ms->comms.data(), n_devices, devices);
        if (r != 0) { ms->comms.clear(); cudaGetLastError(); }
        else ms->exchange = "nccl";
    }
    *out = shell;
    return CVG_OK;
}

int cvg_num_devices(const cvg_ctx* c) { return !c ? 0 : (c->multi ? (int)c->multi->subs.size() : 1); }
const char* cvg_exchange_kind(const cvg_ctx* c) { return (c && c->multi) ? c->multi->exchange : "none"; }

void cvg_destroy(cvg_ctx* c)
{
    if (!c) return;
    if (c->multi) { multi_destroy(c); delete c; return; }
    destroy_lanes(c);
    eng_destroy(c);
}

int cvg_set_lanes(cvg_ctx* c, int n_lanes)
{
    if (!c || n_lanes < 0 || n_lanes > 8) return cvg_set_err(CVG_ERR_INVALID, "cvg_set_lanes: 0 (default) .. 8");
    if (c->multi) { for (cvg_ctx* s : c->multi->subs) s->lanes_cfg = n_lanes ? n_lanes : -1; return CVG_OK; }
    c->lanes_cfg = n_lanes ? n_lanes : -1;
    return CVG_OK;
}

int64_t cvg_launch_count(const cvg_ctx* c)
{
    if (!c) return 0;
    int64_t n = 0;
    if (c->multi) { for (const cvg_ctx* s : c->multi->subs) n += cvg_launch_count(s); return n; }
    n = c->launches;
    for (const Lane* l : c->lanes) n += l->eng->launches;
    return n;
}

// ---- model set ------------------------------------------------------------------------------------------------
int cvg_models_upload(cvg_ctx* c, const float* desc, const float* kpt_xy, const int32_t* view_offsets,
                      const int32_t* view_model, int n_views, cvg_models** out)
{
    if (!c || !out) return cvg_set_err(CVG_ERR_INVALID, "cvg_models_upload: bad argument");
    if (!c->multi) return eng_models_upload(c, desc, kpt_xy, view_offsets, view_model, n_views, out);
    *out = nullptr;
    cvg_models* shell = new cvg_models();
    for (cvg_ctx* s : c->multi->subs) {                     // replicated: 7 MB for the reference's 89 views
        cvg_models* r = nullptr;
        const int rc = eng_models_upload(s, desc, kpt_xy, view_offsets, view_model, n_views, &r);
        if (rc) { for (size_t d = 0; d < shell->replicas.size(); d++) eng_models_free(c->multi->subs[d], shell->replicas[d]); delete shell; return rc; }
        shell->replicas.push_back(r);
    }
    const cvg_models* r0 = shell->replicas[0];
    shell->n_rows = r0->n_rows; shell->n_pad = r0->n_pad; shell->n_views = r0->n_views; shell->view_offsets = r0->view_offsets;
    shell->view_model = r0->view_model; shell->nonint = r0->nonint; shell->max_norm2 = r0->max_norm2; shell->uid = r0->uid;
    *out = shell;
    return CVG_OK;
}

void cvg_models_free(cvg_ctx* c, cvg_models* m)
{
    if (!m) return;
    if (!m->replicas.empty()) {
        for (size_t d = 0; d < m->replicas.size(); d++)
            eng_models_free((c && c->multi && d < c->multi->subs.size()) ? c->multi->subs[d] : nullptr, m->replicas[d]);
        delete m;
        return;
    }
    eng_models_free(c && !c->multi ? c : nullptr, m);
}

// ---- scene batches --------------------------------------------------------------------------------------------
static int scenes_upload_any(cvg_ctx* c, const float* desc, const uint8_t* desc_u8, const float* kpt_xy, const int64_t* offsets,
                             int n_scenes, cvg_scenes** out, bool async)
{
    if (!c || !out || !offsets || n_scenes < 0) return cvg_set_err(CVG_ERR_INVALID, "cvg_scenes_upload: bad argument");
    if (!c->multi) return eng_scenes_upload(c, desc, desc_u8, kpt_xy, offsets, n_scenes, out, async);
    *out = nullptr;
    MultiState* ms = c->multi;
    const size_t nd = ms->subs.size();
    for (int s = 0; s < n_scenes; s++)
        if (offsets[s + 1] < offsets[s]) return cvg_set_err(CVG_ERR_INVALID, "offsets must be non-decreasing");
    cvg_scenes* shell = new cvg_scenes();
    shell->n_scenes_total = n_scenes;
    shell->offsets_copy.assign(offsets, offsets + n_scenes + 1);
    shell->dev_of_scene.assign((size_t)n_scenes, 0); shell->local_of_scene.assign((size_t)n_scenes, 0);
    shell->subs.assign(nd, nullptr);
    std::vector<std::vector<int>> mine(nd);
    if ((size_t)n_scenes < 2 * nd) {
        // a small batch (the five scales of one test image) stays whole on one device, batches take the devices in turn:
        // the caller pipelines them with cvg_detect_scenes_submit
        const size_t d = ms->next_dev++ % nd;
        for (int s = 0; s < n_scenes; s++) mine[d].push_back(s);
    } else {
        // longest-processing-time-first by train rows (the cost of a scene is rows x model rows), ties to the lower device
        std::vector<int> order((size_t)n_scenes);
        for (int s = 0; s < n_scenes; s++) order[(size_t)s] = s;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b]; });
        std::vector<int64_t> load(nd, 0);
        for (int s : order) {
            size_t best = 0;
            for (size_t d = 1; d < nd; d++) if (load[d] < load[best]) best = d;
            mine[best].push_back(s); load[best] += offsets[s + 1] - offsets[s];
        }
        for (auto& v : mine) std::sort(v.begin(), v.end());
    }
    for (size_t d = 0; d < nd; d++) {
        if (mine[d].empty()) continue;
        std::vector<int64_t> off_d(mine[d].size() + 1, 0), src0(mine[d].size());
        for (size_t k = 0; k < mine[d].size(); k++) {
            const int s = mine[d][k];
            shell->dev_of_scene[(size_t)s] = (int)d; shell->local_of_scene[(size_t)s] = (int)k;
            src0[k] = offsets[s]; off_d[k + 1] = off_d[k] + (offsets[s + 1] - offsets[s]);
        }
        const int rc = eng_scenes_upload(ms->subs[d], desc, desc_u8, kpt_xy, off_d.data(), (int)mine[d].size(), &shell->subs[d], async,
                                         src0.data());
        if (rc) { for (size_t e = 0; e < nd; e++) if (shell->subs[e]) eng_scenes_free(ms->subs[e], shell->subs[e]); delete shell; return rc; }
    }
    *out = shell;
    return CVG_OK;
}

int cvg_scenes_upload(cvg_ctx* c, const float* desc, const float* kpt_xy, const int64_t* offsets, int n_scenes, cvg_scenes** out)
{
    return scenes_upload_any(c, desc, nullptr, kpt_xy, offsets, n_scenes, out, false);
}
int cvg_scenes_upload_async(cvg_ctx* c, const float* desc, const float* kpt_xy, const int64_t* offsets, int n_scenes, cvg_scenes** out)
{
    return scenes_upload_any(c, desc, nullptr, kpt_xy, offsets, n_scenes, out, true);
}
int cvg_scenes_upload_u8_async(cvg_ctx* c, const uint8_t* desc, const float* kpt_xy, const int64_t* offsets, int n_scenes, cvg_scenes** out)
{
    return scenes_upload_any(c, nullptr, desc, kpt_xy, offsets, n_scenes, out, true);
}

int cvg_scenes_wait(cvg_ctx* c, cvg_scenes* sc)
{
    if (!c || !sc) return cvg_set_err(CVG_ERR_INVALID, "cvg_scenes_wait: NULL argument");
    if (!c->multi) return eng_scenes_wait(c, sc);
    for (size_t d = 0; d < sc->subs.size(); d++)
        if (sc->subs[d]) { const int rc = eng_scenes_wait(c->multi->subs[d], sc->subs[d]); if (rc) return rc; }
    return CVG_OK;
}

void cvg_scenes_free(cvg_ctx* c, cvg_scenes* sc)
{
    if (!sc) return;
    if (!sc->subs.empty()) {
        for (size_t d = 0; d < sc->subs.size(); d++)
            if (sc->subs[d]) eng_scenes_free((c && c->multi && d < c->multi->subs.size()) ? c->multi->subs[d] : nullptr, sc->subs[d]);
        delete sc;
        return;
    }
    eng_scenes_free(c && !c->multi ? c : nullptr, sc);
}

// ---- fused path over resident scene batches ---------------------------------------------------------------------
int cvg_detect_scenes_submit(cvg_ctx* c, const cvg_models* m, const cvg_scenes* scenes, const float* scales,
                             const cvg_detect_params* p, cvg_pair_result* per_pair, float* inlier_scene_xy,
                             int64_t* inlier_offsets, cvg_job** job)
{
    if (!c || !m || !scenes || !per_pair || !job) return cvg_set_err(CVG_ERR_INVALID, "cvg_detect_scenes_submit: NULL argument");
    const int rc = check_detect_params(p);
    if (rc) return rc;
    if (c->multi) return multi_submit(c, m, scenes, scales, p, per_pair, inlier_scene_xy, inlier_offsets, false, job);
    return dev_submit(c, m, const_cast<cvg_scenes*>(scenes), scales, p, per_pair, inlier_scene_xy, inlier_offsets, false, job);
}

int cvg_job_wait(cvg_ctx* c, cvg_job* job)
{
    if (!c || !job) return cvg_set_err(CVG_ERR_INVALID, "cvg_job_wait: NULL argument");
    const int rc = job->subs.empty() && !job->shell ? dev_wait(job) : multi_wait(job);
    delete job;
    return rc;
}

int cvg_detect_scenes_inliers(cvg_ctx* c, const cvg_models* m, const cvg_scenes* scenes, const float* scales,
                              const cvg_detect_params* p, cvg_pair_result* per_pair, float* inlier_scene_xy,
                              int64_t* inlier_offsets)
{
    if (!c || !m || !scenes || !per_pair) return cvg_set_err(CVG_ERR_INVALID, "cvg_detect_scenes: NULL argument");
    int rc = check_detect_params(p);
    if (rc) return rc;
    cvg_job* job = nullptr;
    if (c->multi) {
        rc = multi_submit(c, m, scenes, scales, p, per_pair, inlier_scene_xy, inlier_offsets, true, &job);
        if (rc) return rc;
        rc = multi_wait(job);
        delete job;
        return rc;
    }
    cvg_scenes* sc = const_cast<cvg_scenes*>(scenes);
    std::vector<std::pair<int, int>> ranges;
    plan_chunks(c, m, sc, true, ranges);
    if (ranges.size() <= 1) {
        // small call: the context's own engine, on the caller's thread
        c->wave_div = 1;
        const bool want_inl = inlier_scene_xy != nullptr && inlier_offsets != nullptr;
        std::vector<float> pool; std::vector<int32_t> cnt;
        rc = eng_detect_range(c, m, sc, 0, sc->ts.n_segs, scales, p, per_pair, want_inl ? &pool : nullptr, want_inl ? &cnt : nullptr);
        if (rc || !want_inl) return rc;
        memcpy(inlier_scene_xy, pool.data(), pool.size() * 4);
        int64_t o = 0;
        for (size_t pair = 0; pair < cnt.size(); pair++) { inlier_offsets[pair] = o; o += cnt[pair]; }
        inlier_offsets[cnt.size()] = o;
        return CVG_OK;
    }
    rc = dev_submit(c, m, sc, scales, p, per_pair, inlier_scene_xy, inlier_offsets, true, &job);
    if (rc) return rc;
    rc = dev_wait(job);
    delete job;
    return rc;
}

int cvg_detect_scenes(cvg_ctx* c, const cvg_models* m, const cvg_scenes* scenes, const float* scales,
                      const cvg_detect_params* p, cvg_pair_result* per_pair)
{
    return cvg_detect_scenes_inliers(c, m, scenes, scales, p, per_pair, nullptr, nullptr);
}

// ---- train-tile sharded match: SURVEY 8e-2, BASELINE config 5 ------------------------------------------------------
// Every device gets the whole query matrix and one contiguous, 256-aligned range of the train rows, computes its local
// top-2 per query with GLOBAL train indices (cvg_dev_match_top2), then ONE exchange: all-gather of the (distance, index)
// pairs, 16 B per query and device, and the lexicographic (distance, index) merge that reproduces OpenCV's tie rule
// whatever the number of shards (cvg_dev_merge_top2) — replaces src/TestsDetector.cpp:59-72 for one huge pair.
int cvg_match_knn2_sharded(cvg_ctx* c, const float* query, int n_query, const float* train, int n_train, float ratio,
                           int32_t* idx, float* dist, uint8_t* accept)
{
    if (!c || !c->multi) return cvg_set_err(CVG_ERR_INVALID, "cvg_match_knn2_sharded needs a context made by cvg_create_multi");
    if (n_query < 0 || n_train < 0 || (n_query > 0 && !query) || (n_train > 0 && !train) || !idx || !dist)
        return cvg_set_err(CVG_ERR_INVALID, "cvg_match_knn2_sharded: bad argument");
    if (n_query == 0) return CVG_OK;
    MultiState* ms = c->multi;
    const int nd = (int)ms->subs.size();
    int64_t per = ((int64_t)n_train + nd - 1) / nd;
    per = (per + TILE_N - 1) / TILE_N * TILE_N;
    std::vector<Done> done((size_t)nd);
    const size_t nq = (size_t)n_query;
    for (int d = 0; d < nd; d++) {
        ms->workers[(size_t)d]->post([=, &done] {
            cvg_ctx* e = ms->subs[(size_t)d];
            const int64_t a = std::min<int64_t>((int64_t)d * per, n_train), b = std::min<int64_t>(a + per, n_train);
            const int rc = [&]() -> int {
                CU_CHECK(cudaSetDevice(e->device));
                CU_CHECK(ms->qbuf[(size_t)d].ensure(nq * DIM * 4));
                CU_CHECK(ms->tbuf[(size_t)d].ensure((size_t)std::max<int64_t>(b - a, 1) * DIM * 4));
                CU_CHECK(ms->loc_d[(size_t)d].ensure(nq * 8)); CU_CHECK(ms->loc_i[(size_t)d].ensure(nq * 8));
                CU_CHECK(ms->all_d[(size_t)d].ensure(nq * 8 * nd)); CU_CHECK(ms->all_i[(size_t)d].ensure(nq * 8 * nd));
                CU_CHECK(cudaMemcpyAsync(ms->qbuf[(size_t)d].p, query, nq * DIM * 4, cudaMemcpyHostToDevice, e->stream));
                if (b > a) CU_CHECK(cudaMemcpyAsync(ms->tbuf[(size_t)d].p, train + (size_t)a * DIM, (size_t)(b - a) * DIM * 4, cudaMemcpyHostToDevice, e->stream));
                return cvg_dev_match_top2(e, e->stream, ms->qbuf[(size_t)d].as<float>(), n_query, ms->tbuf[(size_t)d].as<float>(), (int)(b - a),
                                          (int32_t)a, ms->loc_d[(size_t)d].as<float>(), ms->loc_i[(size_t)d].as<int32_t>());
            }();
            done[(size_t)d].set(rc, cvg_last_error());
        });
    }
    int rc = CVG_OK; std::string err;
    for (int d = 0; d < nd; d++) { const int r = done[(size_t)d].wait(); if (r && !rc) { rc = r; err = done[(size_t)d].err; } }
    if (rc) return cvg_set_err(rc, "%s", err.c_str());
    // ---- the exchange step -------------------------------------------------------------------------------------
    if (!ms->comms.empty()) {
        NcclApi& na = nccl_api();
        int r = na.GroupStart();
        for (int d = 0; d < nd && r == 0; d++) {
            cvg_ctx* e = ms->subs[(size_t)d];
            r = na.AllGather(ms->loc_d[(size_t)d].p, ms->all_d[(size_t)d].p, nq * 8, NCCL_CHAR, ms->comms[(size_t)d], e->stream);
            if (r == 0) r = na.AllGather(ms->loc_i[(size_t)d].p, ms->all_i[(size_t)d].p, nq * 8, NCCL_CHAR, ms->comms[(size_t)d], e->stream);
        }
        const int r2 = na.GroupEnd();
        if (r == 0) r = r2;
        if (r != 0) return cvg_set_err(CVG_ERR_CUDA, "ncclAllGather failed: %s", na.GetErrorString ? na.GetErrorString(r) : "?");
    } else {
        // logical devices on one GPU / no NCCL: the partial results travel to device 0 by device-to-device copies
        cvg_ctx* e0 = ms->subs[0];
        for (int d = 0; d < nd; d++) {
            cvg_ctx* e = ms->subs[(size_t)d];
            CU_CHECK(cudaSetDevice(e->device));
            CU_CHECK(cudaEventRecord(ms->ev[(size_t)d], e->stream));
            CU_CHECK(cudaSetDevice(e0->device));
            CU_CHECK(cudaStreamWaitEvent(e0->stream, ms->ev[(size_t)d], 0));
            CU_CHECK(cudaMemcpyPeerAsync(ms->all_d[0].as<uint8_t>() + (size_t)d * nq * 8, e0->device, ms->loc_d[(size_t)d].p, e->device, nq * 8, e0->stream));
            CU_CHECK(cudaMemcpyPeerAsync(ms->all_i[0].as<uint8_t>() + (size_t)d * nq * 8, e0->device, ms->loc_i[(size_t)d].p, e->device, nq * 8, e0->stream));
        }
    }
    // ---- merge on device 0, results to the host; the other devices only have to drain ----------------------------
    cvg_ctx* e0 = ms->subs[0];
    CU_CHECK(cudaSetDevice(e0->device));
    CU_CHECK(ms->out_idx[0].ensure(nq * 8)); CU_CHECK(ms->out_dist[0].ensure(nq * 8)); CU_CHECK(ms->out_acc[0].ensure(nq));
    rc = cvg_dev_merge_top2(e0, e0->stream, ms->all_d[0].as<float>(), ms->all_i[0].as<int32_t>(), nd, n_query, ratio,
                            ms->out_idx[0].as<int32_t>(), ms->out_dist[0].as<float>(), ms->out_acc[0].as<uint8_t>());
    if (rc) return rc;
    CU_CHECK(cudaMemcpyAsync(idx, ms->out_idx[0].p, nq * 8, cudaMemcpyDeviceToHost, e0->stream));
    CU_CHECK(cudaMemcpyAsync(dist, ms->out_dist[0].p, nq * 8, cudaMemcpyDeviceToHost, e0->stream));
    if (accept) CU_CHECK(cudaMemcpyAsync(accept, ms->out_acc[0].p, nq, cudaMemcpyDeviceToHost, e0->stream));
    for (int d = nd - 1; d >= 0; d--) {
        CU_CHECK(cudaSetDevice(ms->subs[(size_t)d]->device));
        CU_CHECK(cudaStreamSynchronize(ms->subs[(size_t)d]->stream));
    }
    return CVG_OK;
}

}  // extern "C"
