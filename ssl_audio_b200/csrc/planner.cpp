// Host planner: replays, for a whole batch, the random draws the reference makes per sample and
// emits the parameter table consumed by the view kernel.  Bit-exact re-statement of
//   numpy legacy RandomState (MT19937): random_sample / uniform / rand / randint  and
//   CPython `random` (MT19937): randint -> randrange -> _randbelow_with_getrandbits
// in the call order of the reference (paths relative to /root/reference):
//   datasets.py:344  np.random.randint(l - crop_frames)            time crop of a precomputed log-mel
//   datasets.py:112  random.randint(0, len - unit_length)          wav crop
//   augmentations.py:105,108  np.random.random(), np.random.randint(len(bank))      Mixup
//   augmentations.py:136      np.random.rand()                                       MixGaussianNoise (lambd)
//   augmentations.py:34-37    np.random.uniform x2, random.randint x2 (conditional)  RandomResizeCrop
//   augmentations.py:70       np.random.rand(2)                                      RandomLinearFader
// The generator states are imported from / exported to the interpreter's global generators by the
// Python wrapper, so seeding with np.random.seed / random.seed keeps its meaning.
#include "../../include/abt_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace abt {
int set_error(int code, const char* fmt, ...);
}

namespace {

// State regeneration of MT19937.  Both loops only carry anti-dependences (element kk reads kk + 1 before it is rewritten) or
// read values produced 227 elements earlier, so the compiler vectorises them; the clone for AVX2 hosts (picked at load time,
// every x86 B200 host has it) regenerates the state about 1.5x faster than the baseline SSE2 build.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
void mt19937_regenerate(uint32_t* key) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; kk++) {
        y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
        key[kk] = key[kk + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; kk < 623; kk++) {
        y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
        key[kk] = key[kk + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (key[623] & UPPER) | (key[0] & LOWER);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
}

struct MT19937 {
    uint32_t key[624];
    int pos;
    void gen() {
        mt19937_regenerate(key);
        pos = 0;
    }
    uint32_t next32() {
        if (pos >= 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // 53-bit double in [0, 1): numpy random_sample / CPython random()
    double next_double() {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

// numpy legacy RandomState.randint(n) for 0 < n <= 2^32 (masked rejection, 32-bit draws)
uint32_t np_randint(MT19937& g, uint32_t n) {
    const uint32_t rng = n - 1;
    if (rng == 0) return 0;
    if (rng == 0xFFFFFFFFu) return g.next32();
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = g.next32() & mask; } while (v > rng);
    return v;
}

// CPython random._randbelow_with_getrandbits(n), 0 < n < 2^32
uint32_t py_randbelow(MT19937& g, uint32_t n) {
    const int k = 32 - __builtin_clz(n);     // n.bit_length()
    uint32_t r;
    do { r = g.next32() >> (32 - k); } while (r >= n);
    return r;
}

int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

struct abt_planner {
    abt_plan_config cfg;
    MT19937 np_rng;
    MT19937 py_rng;
    std::vector<int64_t> bank;   // circular buffer of clip uids (MixupBYOLA.memory_bank), oldest at bank_head
    int bank_head = 0, bank_count = 0;
    int64_t next_uid;
    // bank_head < cap and idx <= bank_count <= cap: one conditional subtraction instead of a division
    int wrap(int i) const { const int cap = (int)bank.size(); return i >= cap ? i - cap : i; }
    int64_t bank_at(int idx) const { return bank[wrap(bank_head + idx)]; }
    void bank_push(int64_t uid) {
        const int cap = (int)bank.size();
        if (bank_count < cap) { bank[wrap(bank_head + bank_count)] = uid; ++bank_count; }
        else { bank[bank_head] = uid; bank_head = wrap(bank_head + 1); }   // (bank + [x])[-n:]
    }
};

extern "C" int abt_planner_create(const abt_plan_config* cfg, abt_planner** out) {
    if (cfg == nullptr || out == nullptr) return abt::set_error(ABT_ERR_ARG, "null argument");
    if (cfg->n_memory < 1 || cfg->ring_slots < cfg->n_memory) return abt::set_error(ABT_ERR_ARG, "need 1 <= n_memory <= ring_slots");
    if (cfg->n_global < 0 || cfg->n_local < 0 || cfg->n_global + cfg->n_local < 1) return abt::set_error(ABT_ERR_ARG, "need at least one view per clip");
    if (cfg->in_h < 1 || cfg->in_w < 1 || cfg->canvas_h < 1 || cfg->canvas_w < 1) return abt::set_error(ABT_ERR_ARG, "bad geometry");
    abt_planner* p = new abt_planner();
    p->cfg = *cfg;
    std::memset(p->np_rng.key, 0, sizeof(p->np_rng.key));
    std::memset(p->py_rng.key, 0, sizeof(p->py_rng.key));
    p->np_rng.pos = 624;
    p->py_rng.pos = 624;
    p->next_uid = 0;
    p->bank.assign(cfg->n_memory, 0);
    *out = p;
    return 0;
}

extern "C" int abt_planner_destroy(abt_planner* p) {
    delete p;
    return 0;
}

static int set_state(MT19937& g, const uint32_t* key, int pos) {
    if (key == nullptr || pos < 0 || pos > 624) return abt::set_error(ABT_ERR_ARG, "bad MT19937 state (pos %d)", pos);
    std::memcpy(g.key, key, sizeof(g.key));
    g.pos = pos;
    return 0;
}

extern "C" int abt_planner_set_numpy_state(abt_planner* p, const uint32_t* key, int pos) {
    if (p == nullptr) return abt::set_error(ABT_ERR_ARG, "planner is null");
    return set_state(p->np_rng, key, pos);
}
extern "C" int abt_planner_get_numpy_state(const abt_planner* p, uint32_t* key, int* pos) {
    if (p == nullptr || key == nullptr || pos == nullptr) return abt::set_error(ABT_ERR_ARG, "null argument");
    std::memcpy(key, p->np_rng.key, sizeof(p->np_rng.key));
    *pos = p->np_rng.pos;
    return 0;
}
extern "C" int abt_planner_set_pyrandom_state(abt_planner* p, const uint32_t* key, int pos) {
    if (p == nullptr) return abt::set_error(ABT_ERR_ARG, "planner is null");
    return set_state(p->py_rng, key, pos);
}
extern "C" int abt_planner_get_pyrandom_state(const abt_planner* p, uint32_t* key, int* pos) {
    if (p == nullptr || key == nullptr || pos == nullptr) return abt::set_error(ABT_ERR_ARG, "null argument");
    std::memcpy(key, p->py_rng.key, sizeof(p->py_rng.key));
    *pos = p->py_rng.pos;
    return 0;
}

extern "C" int abt_planner_bank_len(const abt_planner* p) { return p ? p->bank_count : 0; }
extern "C" int abt_planner_bank_reset(abt_planner* p) {
    if (p) { p->bank_head = 0; p->bank_count = 0; }
    return 0;
}

static void plan_rrc(abt_planner* p, abt_view_params* v, int out_canvas_h, int out_canvas_w, const double* fs, const double* ts) {
    const abt_plan_config& c = p->cfg;
    // get_params(canvas, in_size, time_scale, freq_scale): h from freq_scale, w from time_scale
    const double uh = fs[0] + (fs[1] - fs[0]) * p->np_rng.next_double();
    const int h = clipi((int)(uh * c.in_h), 1, out_canvas_h);
    const double uw = ts[0] + (ts[1] - ts[0]) * p->np_rng.next_double();
    const int w = clipi((int)(uw * c.in_w), 1, out_canvas_w);
    const int i = (out_canvas_h > h) ? (int)py_randbelow(p->py_rng, (uint32_t)(out_canvas_h - h + 1)) : 0;
    const int j = (out_canvas_w > w) ? (int)py_randbelow(p->py_rng, (uint32_t)(out_canvas_w - w + 1)) : 0;
    v->i = i; v->j = j; v->h = h; v->w = w;
    v->flags |= 2;
}

extern "C" int abt_planner_plan_batch(abt_planner* p, int n_clips, int time_crop_range, int wav_crop_range, int32_t* starts,
                                      int32_t* wav_starts, abt_view_params* params, int32_t* slots) {
    if (p == nullptr || params == nullptr) return abt::set_error(ABT_ERR_ARG, "null argument");
    if (n_clips < 0) return abt::set_error(ABT_ERR_ARG, "negative n_clips");
    const abt_plan_config& c = p->cfg;
    if (c.mixup && n_clips > c.ring_slots - c.n_memory && p->cfg.n_global > 0)
        return abt::set_error(ABT_ERR_ARG, "batch of %d clips needs ring_slots >= n_memory + batch (have %d)", n_clips, c.ring_slots);
    const int n_views = c.n_global + c.n_local;
    const int64_t first_uid = p->next_uid;
    for (int b = 0; b < n_clips; ++b) {
        const int64_t uid = p->next_uid++;
        if (slots) slots[b] = (int32_t)(uid % c.ring_slots);
        if (wav_starts) wav_starts[b] = -1;
        if (wav_crop_range > 0) {   // random.randint(0, range)
            const int s = (int)py_randbelow(p->py_rng, (uint32_t)wav_crop_range + 1u);
            if (wav_starts) wav_starts[b] = s;
        }
        if (starts) starts[b] = -1;
        if (time_crop_range > 0) {  // np.random.randint(range)
            const int s = (int)np_randint(p->np_rng, (uint32_t)time_crop_range);
            if (starts) starts[b] = s;
        }
        for (int v = 0; v < n_views; ++v) {
            abt_view_params* vp = &params[(size_t)b * n_views + v];
            std::memset(vp, 0, sizeof(*vp));
            vp->out_index = v;
            vp->w_x = 1.0f;
            if (v < c.n_global) {
                if (c.mixup) {
                    vp->flags |= 1;
                    const double alpha = (double)c.mixup_ratio_d * p->np_rng.next_double();
                    if (p->bank_count > 0) {
                        const uint32_t idx = np_randint(p->np_rng, (uint32_t)p->bank_count);
                        const int64_t zuid = p->bank_at((int)idx);
                        if (zuid >= first_uid) { vp->z_kind = 2; vp->z_index = (int32_t)(zuid - first_uid); }
                        else { vp->z_kind = 1; vp->z_index = (int32_t)(zuid % c.ring_slots); }
                        // log_mixup_exp(x, z, 1. - alpha): weights float32(a), float32(1. - a), a = 1. - alpha
                        const double a = 1.0 - alpha;
                        vp->w_x = (float)a;
                        vp->w_z = (float)(1.0 - a);
                    }
                    p->bank_push(uid);
                }
                if (c.gnoise) {
                    const double lambd = (double)c.gnoise_ratio_d * p->np_rng.next_double();
                    vp->g_lambda = (float)lambd;
                    vp->g_keep = (float)(1.0 - lambd);
                    vp->flags |= 8;
                }
                if (c.rrc) plan_rrc(p, vp, c.canvas_h, c.canvas_w, c.freq_scale, c.time_scale);
                if (c.rlf) {
                    const double u0 = p->np_rng.next_double(), u1 = p->np_rng.next_double();
                    vp->head = (float)(c.fader_gain * (2.0 * u0 - 1.0));
                    vp->tail = (float)(c.fader_gain * (2.0 * u1 - 1.0));
                    vp->flags |= 4;
                }
            } else {
                // local crop: RandomResizeCrop(local size, virtual_crop_scale=(1,1), scale=local_scale)
                plan_rrc(p, vp, c.in_h, c.in_w, c.local_scale, c.local_scale);
            }
        }
    }
    return 0;
}

static size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }

extern "C" int abt_planner_packed_bytes(const abt_planner* p, int n_clips, size_t* total, size_t* off_starts, size_t* off_wav_starts,
                                        size_t* off_slots) {
    if (p == nullptr || n_clips < 0) return abt::set_error(ABT_ERR_ARG, "bad argument");
    const size_t n_views = (size_t)(p->cfg.n_global + p->cfg.n_local);
    const size_t o1 = a16(sizeof(abt_view_params) * n_views * (size_t)n_clips);
    const size_t o2 = o1 + a16(sizeof(int32_t) * (size_t)n_clips);
    const size_t o3 = o2 + a16(sizeof(int32_t) * (size_t)n_clips);
    if (off_starts) *off_starts = o1;
    if (off_wav_starts) *off_wav_starts = o2;
    if (off_slots) *off_slots = o3;
    if (total) *total = o3 + a16(sizeof(int32_t) * (size_t)n_clips);
    return 0;
}

extern "C" int abt_planner_plan_batch_packed(abt_planner* p, int n_clips, int time_crop_range, int wav_crop_range, void* out, size_t out_bytes) {
    size_t total, o1, o2, o3;
    if (int rc = abt_planner_packed_bytes(p, n_clips, &total, &o1, &o2, &o3)) return rc;
    if (out == nullptr || out_bytes < total) return abt::set_error(ABT_ERR_ARG, "packed plan buffer too small (%zu < %zu)", out_bytes, total);
    uint8_t* b = static_cast<uint8_t*>(out);
    return abt_planner_plan_batch(p, n_clips, time_crop_range, wav_crop_range, reinterpret_cast<int32_t*>(b + o1),
                                  reinterpret_cast<int32_t*>(b + o2), reinterpret_cast<abt_view_params*>(b), reinterpret_cast<int32_t*>(b + o3));
}

// Plan one batch against the INTERPRETER's global generators in place: np_state points at numpy's legacy
// `struct { uint32_t key[624]; int pos; }` (may be NULL if no numpy draw can happen), py_index / py_key at the index and the
// 624 state words of CPython's `random` instance (may be NULL likewise).  One call instead of set-state / plan / get-state.
extern "C" int abt_planner_plan_batch_global(abt_planner* p, int n_clips, int time_crop_range, int wav_crop_range, void* np_state,
                                             int32_t* py_index, uint32_t* py_key, void* out, size_t out_bytes) {
    if (p == nullptr) return abt::set_error(ABT_ERR_ARG, "planner is null");
    struct NpState { uint32_t key[624]; int pos; };
    NpState* ns = static_cast<NpState*>(np_state);
    if (ns != nullptr) {
        if (ns->pos < 0 || ns->pos > 624) return abt::set_error(ABT_ERR_ARG, "bad numpy MT19937 position %d", ns->pos);
        std::memcpy(p->np_rng.key, ns->key, sizeof(ns->key));
        p->np_rng.pos = ns->pos;
    }
    if (py_index != nullptr && py_key != nullptr) {
        if (*py_index < 0 || *py_index > 624) return abt::set_error(ABT_ERR_ARG, "bad CPython MT19937 index %d", *py_index);
        std::memcpy(p->py_rng.key, py_key, sizeof(p->py_rng.key));
        p->py_rng.pos = *py_index;
    }
    const int rc = abt_planner_plan_batch_packed(p, n_clips, time_crop_range, wav_crop_range, out, out_bytes);
    if (rc != 0) return rc;
    if (ns != nullptr) {
        std::memcpy(ns->key, p->np_rng.key, sizeof(ns->key));
        ns->pos = p->np_rng.pos;
    }
    if (py_index != nullptr && py_key != nullptr) {
        std::memcpy(py_key, p->py_rng.key, sizeof(p->py_rng.key));
        *py_index = p->py_rng.pos;
    }
    return 0;
}
