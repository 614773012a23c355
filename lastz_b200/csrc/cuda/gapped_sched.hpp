/*
 * gapped_sched.hpp -- the anchor loop of the gapped stage (gapped_extend gapped_extend.c:1012, :1300-1470), host C++.
 *
 * The reference extends anchors best score first, one at a time; every alignment it keeps constrains the later
 * ones (msp_left_right :3953, get_above_below :4043, update_LR_bounds :4588, update_active_segs :4885).  Here the
 * one-sided sweeps of MANY anchors run at once on the device and the results are committed strictly in the
 * reference's order:
 *
 *   - an anchor starts as soon as a lane (a set of device buffers) is free, against the alignments committed so
 *     far ("snapshot"), unless it sits where an earlier, still unresolved anchor's alignment will probably run --
 *     such anchors are usually skipped by the reference (`continue` at :1330) and wait for that anchor;
 *   - an anchor is committed when every earlier anchor is resolved and both of its sweeps are VALID against all
 *     committed alignments.  An alignment X committed after a sweep's snapshot can reach the sweep only in three
 *     ways: (1) the anchor lies on X -- the anchor is skipped; (2) X crosses the anchor's row -- it may be a new
 *     left/right neighbour: if msp_left_right now answers differently both sweeps restart; (3) X lies wholly above
 *     (below) the anchor's row -- it enters the forward (reverse) sweep's aboveList (belowList) and is first looked at
 *     in the sweep row where X begins (ends).  Rows before that one are untouched, so the sweep is RESUMED from its
 *     last checkpoint before that row (k_ydrop_mw writes one every ckptEvery rows) instead of being redone;
 *   - a sweep X cannot reach at all (it ended before that row) is valid as it stands.
 *
 * Every committed result is therefore the result of a sweep that saw exactly the alignments the reference's sweep
 * sees, row for row, and the output equals the sequential algorithm's by construction -- no estimate takes part in
 * correctness; estimates only decide what is worth starting.
 *
 * The scheduler is written against a small backend interface (device buffers, launches, completion polling) so that
 * the SAME code runs on the GPU (gapped.cu) and on the host block emulator (tests/warp_emu/test_gapped_sched.cpp),
 * where it is checked against the oracle without a GPU.
 */
#ifndef LZB_GAPPED_SCHED_HPP
#define LZB_GAPPED_SCHED_HPP

#include <algorithm>
#include <atomic>
#include <chrono>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <unordered_map>
#include <vector>

struct hseg { int type; u32 b1, b2, e1, e2; };
static const segref NOSEG = { -1, -1 };

struct galn {                          /* galign gapped_extend.c:214-245 */
    u32 pos1, pos2, end1, end2; u64 hspId;
    std::vector<hseg> segs;
    segref left1, right1, left2, right2;
    lzb_alignel* align;
    int next, prev;
    int devIx;                         /* index in the device alignment table once committed */
};

#define MAX_RPT ((1u << 30) - 1)
static lzb_editscript* es_new(u32 cap) {
    if (cap < 16) cap = 16;
    lzb_editscript* s = (lzb_editscript*)calloc(1, sizeof(lzb_editscript) + (size_t)(cap - 1) * 4);
    s->size = cap; return s;
}
/* edit_script_add edit_script.c:261 applied to an already run-length-encoded op */
static void es_add(lzb_editscript** ps, u32 op, u32 rpt) {
    lzb_editscript* s = *ps;
    if (s->len > 0 && (s->tailOp & 3) == op) {
        u32 tr = s->op[s->len - 1] >> 2;
        if ((u64)tr + rpt <= MAX_RPT) { s->op[s->len - 1] += rpt << 2; return; }
        s->op[s->len - 1] = op | (MAX_RPT << 2); rpt = tr + rpt - MAX_RPT;
    }
    if (s->len + 2 > s->size) {
        u32 nsz = s->size * 2 + 16;
        s = (lzb_editscript*)realloc(s, sizeof(lzb_editscript) + (size_t)(nsz - 1) * 4); s->size = nsz; *ps = s;
    }
    while (rpt > MAX_RPT) { s->op[s->len++] = op | (MAX_RPT << 2); rpt -= MAX_RPT; }
    s->op[s->len++] = op | (rpt << 2); s->tailOp = op;
}

struct dp_result { s32 score; u32 end1, end2, rows; int status; unsigned long long cells; std::vector<u32> ops; };

/* what the scheduler needs to know about the two sequences and the scoring */
struct gx_input {
    const u8* h_seq1; const u8* h_seq2; u32 len1, len2;
    const s32* hostSub; s32 gapOpen, gapExtend;
};

struct gx {                             /* state of one lzb_gapped_extend call */
    gx_input in;
    const lzb_gapped_params* P;
    std::vector<galn> al; int obi, oed;
    std::vector<int> committed;         /* host alignment indices in commit order = device table order */
    std::vector<dseg> hsegs;            /* device segment table (host mirror, append-only) */
    std::vector<dalign> haligns;        /* device alignment table (host mirror, append-only) */
    lzb_gapped_stats st;
};

/* msp_left_right gapped_extend.c:3953-4040 */
static bool anchor_neighbours(gx& G, galn& m, int* coverer = NULL) {
    u32 pos1 = m.pos1, pos2 = m.pos2, right = 0xFFFFFFFFu, left = 0xFFFFFFFFu;
    segref R = NOSEG, Lf = NOSEG;
    for (int o = G.obi; o >= 0 && G.al[o].pos1 <= pos1; o = G.al[o].next) {
        galn& x = G.al[o];
        if (x.end1 < pos1) continue;
        /* first segment whose e1 >= pos1 (e1 never decreases along an alignment) */
        int ns = (int)x.segs.size(), k = 0, hi2 = ns;
        while (k < hi2) { int mid = (k + hi2) >> 1; if (x.segs[mid].e1 < pos1) k = mid + 1; else hi2 = mid; }
        if (k == ns) continue;
        hseg& bp = x.segs[k]; s32 d;
        if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - pos2) + (s32)(pos1 - bp.b1); else d = (s32)(bp.b2 - pos2);
        if (d == 0) { if (coverer) *coverer = o; return false; }
        if (d > 0 && (u32)d < right) { right = (u32)d; R.al = o; R.sg = k; }
        else if (d < 0 && (u32)-d < left) { left = (u32)-d; Lf.al = o; Lf.sg = k; }
    }
    m.right1 = m.right2 = R; m.left1 = m.left2 = Lf;
    return true;
}

/* align_left_right gapped_extend.c:4078-4175 */
static void alignment_neighbours(gx& G, galn& m) {
    u32 pos1 = m.pos1, pos2 = m.pos2, end1 = m.end1, end2 = m.end2;
    u32 rB = 0xFFFFFFFFu, rT = rB, lB = rB, lT = rB;
    segref RB = NOSEG, RT = NOSEG, LB = NOSEG, LT = NOSEG;
    for (int o = G.obi; o >= 0; o = G.al[o].next) {
        galn& x = G.al[o];
        if (x.pos1 > end1 || x.end1 < pos1) continue;
        int k = 0, ns = (int)x.segs.size();
        while (k < ns && !(x.segs[k].type != SEG_HORZ && x.segs[k].e1 >= pos1)) k++;
        if (k < ns && x.segs[k].b1 <= pos1) {
            hseg& bp = x.segs[k]; s32 d;
            if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - pos2) + (s32)(pos1 - bp.b1); else d = (s32)(bp.b2 - pos2);
            if (d > 0 && (u32)d < rB) { rB = (u32)d; RB.al = o; RB.sg = k; }
            else if (d < 0 && (u32)-d < lB) { lB = (u32)-d; LB.al = o; LB.sg = k; }
        }
        while (k < ns && !(x.segs[k].type != SEG_HORZ && x.segs[k].e1 >= end1)) k++;
        if (k < ns) {
            hseg& bp = x.segs[k]; s32 d;
            if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - end2) + (s32)(end1 - bp.b1); else d = (s32)(bp.b2 - end2);
            if (d > 0 && (u32)d < rT) { rT = (u32)d; RT.al = o; RT.sg = k; }
            else if (d < 0 && (u32)-d < lT) { lT = (u32)-d; LT.al = o; LT.sg = k; }
        }
    }
    m.right1 = RB; m.right2 = RT; m.left1 = LB; m.left2 = LT;
}

/* insert_align gapped_extend.c:4210-4240 */
static void list_insert(gx& G, int mi) {
    galn& m = G.al[mi];
    int qq = -1, p = G.obi;
    while (p >= 0 && G.al[p].pos1 < m.pos1) { qq = p; p = G.al[p].next; }
    if (qq >= 0) { G.al[qq].next = mi; m.next = p; } else { m.next = G.obi; G.obi = mi; }
    qq = -1; p = G.oed;
    while (p >= 0 && G.al[p].end1 > m.end1) { qq = p; p = G.al[p].prev; }
    if (qq >= 0) { G.al[qq].prev = mi; m.prev = p; } else { m.prev = G.oed; G.oed = mi; }
}

/* save_seg gapped_extend.c:5220-5262 */
static void add_diag(galn& m, u32 b1, u32 b2, u32 e1, u32 e2) {
    if (!m.segs.empty()) {
        hseg& last = m.segs.back();
        hseg g; g.type = (b1 == last.e1 + 1) ? SEG_HORZ : SEG_VERT;
        g.b1 = last.e1 + 1; g.b2 = last.e2 + 1; g.e1 = b1 - 1; g.e2 = b2 - 1;
        m.segs.push_back(g);
    }
    hseg d = { SEG_DIAG, b1, b2, e1, e2 };
    m.segs.push_back(d);
}

/* score_alignment gapped_extend.c:5631-5690 (host bytes; O(alignment length), only after lopping) */
static s32 rescore(gx& G, u32 p1, u32 p2, lzb_editscript* s) {
    const s32* sub = G.in.hostSub; s32 sim = 0;
    const u8* s1 = G.in.h_seq1; const u8* s2 = G.in.h_seq2;
    for (u32 k = 0; k < s->len; k++) {
        u32 rpt = s->op[k] >> 2, op = s->op[k] & 3;
        if (!rpt) continue;
        if (op == LZB_OP_SUB) { for (u32 j = 0; j < rpt; j++) sim += sub[(u32)s1[p1 + j] * 256 + s2[p2 + j]]; p1 += rpt; p2 += rpt; }
        else if (op == LZB_OP_INS) { sim -= G.in.gapOpen + (s32)rpt * G.in.gapExtend; p2 += rpt; }
        else { sim -= G.in.gapOpen + (s32)rpt * G.in.gapExtend; p1 += rpt; }
    }
    return sim;
}

static segref dev_ref(gx& G, segref r) { segref o = NOSEG; if (r.al >= 0) { o.al = G.al[r.al].devIx; o.sg = r.sg; } return o; }

/* ---- lanes ---- */
enum { SIDE_IDLE = 0, SIDE_RUNNING = 1, SIDE_DONE = 2, SIDE_PAUSED = 3 };
struct gx_side {
    int phase;
    size_t snapshot;                    /* committed.size() the sweep (or its valid prefix) was computed against */
    int mode;                           /* kernel: 0 four-warp register kernel, 1 one-warp kernel, 2/3 shared-memory ring */
    u32 token;
    bool tbOnly;                        /* the running launch only redoes the traceback */
    dp_result res;                      /* valid when DONE */
    u32 ckptCount, ckptEvery;           /* checkpoints the finished sweep left behind */
    u32 prog0Rows, prog0Used;           /* first progress report seen from the running sweep */
    u64 pausedOn;                       /* PAUSED: the earlier anchor whose alignment the sweep stopped short of */
};
struct gx_lane_state {
    bool busy; u64 anchor; segref left1, right1;   /* the anchor's neighbours the sweeps were started with */
    bool unsure;                        /* started from the edge of another anchor's reach: may yet be skipped */
    gx_side s[2];                       /* 0 = reverse (left) sweep, 1 = forward (right) sweep */
};

/*
 * Backend concept (all calls from the one host thread that runs the scheduler):
 *   int      lanes(int want)                      make sure `want` lanes exist; returns how many do (>= 1) or -1
 *   dp_job*  job(int lane, int side)              host-writable, device-visible descriptor; buffer fields are the backend's
 *   int*     list(int lane, int side, size_t n)   host-writable, device-visible int array of >= n entries (NULL on failure)
 *   int      tables(const dseg* s, size_t s0, size_t s1, const dalign* a, size_t a0, size_t a1)   append s[s0..s1), a[a0..a1)
 *   const dseg* segs();  const dalign* aligns()   device addresses of the tables
 *   int      launch(int mode, const u16* ix, int n)   run jobs ix[k] = 2*lane+side with kernel `mode`
 *   bool     poll()                               give the device a chance / wait a little; false on device failure
 *   const u32* ops(int lane, int side)            host-readable ops of a finished job
 *   int      grow(int lane, int side, int what)   what: DP_TBROW, DP_OPS, DP_ACT -- enlarge that buffer
 *   u32      ckpt_every(), ring(int mode)
 *   const char* error()
 */
template <class Backend>
static int gx_run(Backend& B, const gx_input& in, lzb_segment* anchors, uint64_t n, const lzb_gapped_params* P,
                  lzb_alignel** list, lzb_gapped_stats* stats, int (*fail)(const char*, ...)) {
    auto wall0 = std::chrono::steady_clock::now();
    auto now = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count(); };
    *list = NULL;
    gx G; G.in = in; G.P = P; G.obi = G.oed = -1;
    memset(&G.st, 0, sizeof G.st); G.st.anchors = n;
    const u32 tbLen = 1 + (P->tracebackBytes - 8);                    /* new_traceback :2272-2290 */
    const u32 len1 = in.len1, len2 = in.len2;
    const u8* h1 = in.h_seq1; const u8* h2 = in.h_seq2;
    /* LZB_GAP_TRACE: the scheduler's decisions, collected in memory and written to stderr when the call ends (writing as it
     * goes would slow the very loop it describes) */
    const bool trace = getenv("LZB_GAP_TRACE") != NULL;
    std::string traceBuf;
#define GX_TRACE(...) do { if (trace) { char b_[512]; int n_ = snprintf(b_, sizeof b_, __VA_ARGS__); traceBuf.append(b_, (size_t)(n_ < 511 ? n_ : 511)); } } while (0)
    const bool prof = getenv("LZB_GAP_PROFILE") != NULL;

    /* [multi] sequences (NUL-separated partitions, sequences.h:188-191): a sweep ends at the NULs around its anchor
     * (gapped_extend.c:1357-1372); the limits go into the job's M and N.  Alignments of different partitions cannot
     * meet, so one pass over all anchors gives what the reference's per-partition batches give (gapped_extend.c:1058). */
    std::vector<u32> tSeparators, qSeparators;                   /* positions of the NULs inside each sequence, ascending */
    for (const u8* z = (const u8*)memchr(h1, 0, len1); z; z = (const u8*)memchr(z + 1, 0, len1 - (size_t)(z + 1 - h1)))
        tSeparators.push_back((u32)(z - h1));
    for (const u8* z = (const u8*)memchr(h2, 0, len2); z; z = (const u8*)memchr(z + 1, 0, len2 - (size_t)(z + 1 - h2)))
        qSeparators.push_back((u32)(z - h2));

    /* qSegmentsByDecreasingScore segment.c:1748-1771 (a total order, so any sort gives the same result) */
    std::sort(anchors, anchors + n, [](const lzb_segment& a, const lzb_segment& b) {
        if (a.s != b.s) return a.s > b.s;
        if (a.length != b.length) return a.length < b.length;
        if (a.pos2 != b.pos2) return a.pos2 < b.pos2;
        if (a.pos1 != b.pos1) return a.pos1 < b.pos1;
        return a.id < b.id;
    });
    G.al.resize(n + 1);
    for (u64 i = 0; i <= n; i++) { galn& m = G.al[i]; m.align = NULL; m.next = m.prev = -1; m.devIx = -1; m.left1 = m.right1 = m.left2 = m.right2 = NOSEG; m.pos1 = m.pos2 = m.end1 = m.end2 = 0; m.hspId = 0; }
    for (u64 i = 0; i < n; i++) { G.al[i].pos1 = anchors[i].pos1; G.al[i].pos2 = anchors[i].pos2; G.al[i].hspId = anchors[i].hspId; }

    auto trivial_alignment = [&](u32 pos1, u32 end1, u32 count, s32 s) {
        galn& m = G.al[n];
        m.pos1 = pos1; m.pos2 = 0; m.end1 = end1; m.end2 = count - 1;
        add_diag(m, m.pos1, m.pos2, m.end1, m.end2);
        list_insert(G, (int)n);
        m.devIx = (int)G.committed.size(); G.committed.push_back((int)n);
        lzb_alignel* a = (lzb_alignel*)calloc(1, sizeof *a);
        a->script = es_new(4); es_add(&a->script, LZB_OP_SUB, count);
        a->beg1 = pos1 + 1; a->beg2 = 1; a->end1 = end1 + 1; a->end2 = count; a->seq1 = h1; a->seq2 = h2;
        a->s = s < P->scoreThreshold ? P->scoreThreshold : s; a->isTrivial = 1; m.align = a;
    };
    auto same_bases = [&](u32 off1, u32 count, s32* score) -> bool {
        s32 s = 0; const s32* sub = in.hostSub;
        for (u32 i = 0; i < count; i++) {
            u8 a = h1[off1 + i], b = h2[i];
            if (a >= 'a' && a <= 'z') a -= 32;
            if (b >= 'a' && b <= 'z') b -= 32;
            if (a != b) return false;
            s32 v = sub[(u32)a * 256 + b];
            if (s == 0x7FFFFFFF) ; else if (v <= 0 || s < 0x7FFFFFFF - v) s += v; else s = 0x7FFFFFFF;
        }
        *score = s; return true;
    };
    /* identical_sequences gapped_extend.c:1886-1930 -> trivial self alignment :1113-1151 */
    if (P->identityCheck && len1 == len2) { s32 s; if (same_bases(0, len1, &s)) trivial_alignment(0, len1 - 1, len1, s); }
    /* identical_partition_of_sequence gapped_extend.c:2034-2120 -> :1185-1230: an unpartitioned query that equals
     * one partition of a [multi] target (the first such partition) gets the trivial alignment of that partition */
    if (P->identityCheck && !G.al[n].align && !tSeparators.empty() && qSeparators.empty()) {
        for (size_t k = 0; k < tSeparators.size(); k++) {        /* partition k lies between separator k and the next one (or the end) */
            const u32 before = tSeparators[k], after = k + 1 < tSeparators.size() ? tSeparators[k + 1] : len1;
            s32 s;
            if (after - (before + 1) != len2 || !same_bases(before + 1, len2, &s)) continue;
            trivial_alignment(before + 1, after - 1, len2, s);
            break;
        }
    }

    /* ---- lanes ---- */
    int W = P->speculation; if (W < 1) W = 1; if (W > LZB_LAUNCH_MAX / 2) W = LZB_LAUNCH_MAX / 2;
    { const char* e = getenv("LZB_SPECULATION"); if (e) { W = atoi(e); if (W < 1) W = 1; if (W > LZB_LAUNCH_MAX / 2) W = LZB_LAUNCH_MAX / 2; } }
    if ((u64)W > n) W = n ? (int)n : 1;
    /* lanes are made as they are needed (a lane is 2 x the traceback size of device memory) */
    int have = B.lanes(std::min(W, 8));
    if (have < 1) return -1;
    if (have < std::min(W, 8)) W = have;
    /* kernel a sweep starts with: 1 = the one-warp kernel (1.8 us a row however many sweeps share the device), 0 = the four-warp
     * kernel (1.45 us a row alone, 2.2 with three sweeps per SM), 2/3 = the shared-memory kernel for bands wider than 512 columns */
    int firstMode = 1;
    { const char* e = getenv("LZB_DP_MODE"); if (e) { int mdv = atoi(e); if (mdv >= 0 && mdv <= 3) firstMode = mdv; } }
    std::vector<gx_lane_state> lanes(W);
    for (auto& ln : lanes) { ln.busy = false; ln.s[0].phase = ln.s[1].phase = SIDE_IDLE; }
    std::vector<int> laneOf(n, -1);
    std::vector<u8> fin(n, 0);                               /* 1 = skipped / committed / dropped */
    u32 tokenCounter = 0;
    u64 pfPasses = 0; double pfCommit = 0, pfPoll = 0, pfStart = 0, pfValidate = 0; u64 pfPolls = 0, pfLaunches = 0, pfJobs = 0, pfResumes = 0, pfRestarts = 0, pfWasted = 0;

    /* the anchor POINTS (commit overwrites al[i].pos1/pos2 with the alignment's start) and the anchors by
     * seq-1 position: when an alignment is committed the anchors lying on it are retired there and then
     * (msp_left_right's d == 0 test, gapped_extend.c:4008) */
    std::vector<u32> apos1(n), apos2(n);
    for (u64 i = 0; i < n; i++) { apos1[i] = G.al[i].pos1; apos2[i] = G.al[i].pos2; }
    std::vector<u32> byPos(n);
    for (u64 i = 0; i < n; i++) byPos[i] = (u32)i;
    std::sort(byPos.begin(), byPos.end(), [&](u32 a, u32 b) { return apos1[a] != apos1[b] ? apos1[a] < apos1[b] : a < b; });

    /* device tables: append the alignments committed since the last call */
    size_t segsUp = 0, alignsUp = 0;
    auto push_tables = [&]() -> int {
        for (size_t k = G.haligns.size(); k < G.committed.size(); k++) {
            galn& m = G.al[G.committed[k]];
            dalign d; memset(&d, 0, sizeof d);
            d.pos1 = m.pos1; d.end1 = m.end1;
            d.segBegin = (int)G.hsegs.size(); d.segCount = (int)m.segs.size();
            d.left1 = dev_ref(G, m.left1); d.right1 = dev_ref(G, m.right1); d.left2 = dev_ref(G, m.left2); d.right2 = dev_ref(G, m.right2);
            for (auto& s : m.segs) { dseg x = { s.b1, s.b2, s.e1, s.e2, s.type }; G.hsegs.push_back(x); }
            G.haligns.push_back(d);
        }
        if (G.hsegs.size() > segsUp || G.haligns.size() > alignsUp) {
            if (B.tables(G.hsegs.data(), segsUp, G.hsegs.size(), G.haligns.data(), alignsUp, G.haligns.size())) return -1;
            segsUp = G.hsegs.size(); alignsUp = G.haligns.size();
        }
        return 0;
    };

    /* ---- launches are collected and sent in batches, one per kernel mode ---- */
    std::vector<u16> batch[4];
    auto flush = [&]() -> int {
        for (int md = 0; md < 4; md++) {
            size_t at = 0;
            while (at < batch[md].size()) {
                const int k = (int)std::min<size_t>(LZB_LAUNCH_MAX, batch[md].size() - at);
                if (B.launch(md, batch[md].data() + at, k)) return -1;
                pfLaunches++; pfJobs += (u64)k; at += (size_t)k;
            }
            batch[md].clear();
        }
        return 0;
    };

    const double slackRows = getenv("LZB_GAP_SLACK") ? atof(getenv("LZB_GAP_SLACK")) : 300.0;   /* margin around an expected reach, rows (tests: negative = start everything) */
    const bool rowLimits = !(getenv("LZB_GAP_ROWLIMIT") && !atoi(getenv("LZB_GAP_ROWLIMIT")));
    double reachShared = 0;                                  /* the scheduler's current estimate of a sweep's length (0: none yet) */
    auto reachKnown = [&]() { return reachShared > 0; };
    auto reachNow = [&]() { return reachShared; };
    /* queue one sweep of lane z: fresh (rec < 0), resumed from checkpoint record rec, or traceback only */
    auto queue_side = [&](int z, int side, int rec, bool tbOnly) -> int {
        gx_lane_state& ln = lanes[z]; gx_side& sd = ln.s[side];
        galn& m = G.al[ln.anchor];
        const u32 aPos1 = apos1[ln.anchor], aPos2 = apos2[ln.anchor];
        if (push_tables()) return -1;
        dp_job& J = *B.job(z, side);
        const int rev = side == 0;
        if (!tbOnly) {
            if (rec >= 0 && sd.mode == 1) sd.mode = 0;        /* a sweep that is continued is about to meet an alignment: the wider window */
            J.reversed = rev; J.a1 = aPos1; J.a2 = aPos2;
            u32 low1 = 0, high1 = len1, low2 = 0, high2 = len2;  /* the anchor's partition in each sequence: first base, one past the last */
            if (!tSeparators.empty()) {
                auto after = std::upper_bound(tSeparators.begin(), tSeparators.end(), aPos1);
                if (after != tSeparators.end()) high1 = *after;
                if (after != tSeparators.begin()) low1 = *(after - 1) + 1;
            }
            if (!qSeparators.empty()) {
                auto after = std::upper_bound(qSeparators.begin(), qSeparators.end(), aPos2);
                if (after != qSeparators.end()) high2 = *after;
                if (after != qSeparators.begin()) low2 = *(after - 1) + 1;
            }
            J.M = rev ? aPos1 + 1 - low1 : high1 - (aPos1 + 1); J.N = rev ? aPos2 + 1 - low2 : high2 - (aPos2 + 1);
            /* initial L/R, gapped_extend.c:3500-3543 */
            const segref mLeft = ln.left1, mRight = ln.right1;
            s32 L = 0, R = (s32)(J.N + 1);
            if (mLeft.al >= 0) { hseg& s = G.al[mLeft.al].segs[mLeft.sg]; L = (s32)(s.b2 - aPos2); if (s.type == SEG_DIAG) L -= (s32)(s.b1 - aPos1); }
            if (mRight.al >= 0) { hseg& s = G.al[mRight.al].segs[mRight.sg]; R = (s32)(s.b2 - aPos2); if (s.type == SEG_DIAG) R -= (s32)(s.b1 - aPos1); }
            if (rev) {
                if (mLeft.al < 0 && mRight.al >= 0) { L = -R + 1; R = (s32)(J.N + 1); }
                else if (mLeft.al >= 0 && mRight.al < 0) { R = -L - 1; L = 0; }
                else if (mLeft.al >= 0 && mRight.al >= 0) { s32 tt = -L - 1; L = -R + 1; R = tt; }
            }
            J.L0 = L; J.R0 = R;
            J.leftSeg = dev_ref(G, mLeft); J.rightSeg = dev_ref(G, mRight);
            /* get_above_below :4043-4060, from sweep row `fromRow` on: the alignments that begin above the anchor's row
             * in increasing pos1 (forward sweep), those that end below it in decreasing end1 (reverse sweep) */
            const u32 every = B.ckpt_every();
            const u32 fromRow = rec >= 0 ? ((u32)rec + 1) * every : 0;      /* rows <= fromRow are already done */
            std::vector<int> lst;
            if (!rev) { for (int o = G.obi; o >= 0; o = G.al[o].next) if (G.al[o].pos1 > aPos1 && G.al[o].pos1 - aPos1 > fromRow) lst.push_back(G.al[o].devIx); }
            else { for (int o = G.oed; o >= 0; o = G.al[o].prev) if (G.al[o].end1 < aPos1 && aPos1 - G.al[o].end1 > fromRow) lst.push_back(G.al[o].devIx); }
            int* lv = B.list(z, side, lst.size() + 1);
            if (!lv) return -1;
            for (size_t k = 0; k < lst.size(); k++) lv[k] = lst[k];
            lv[lst.size()] = -1;
            J.listv = lv; J.alignList = 0;
            J.tbLen = tbLen;
            J.resume = rec;
            if (rec < 0) { J.progRows = 0; J.progUsed = 0; }
            /* A sweep towards an EARLIER anchor that is still open will be cut short by that anchor's alignment.  Stop it a little
             * beyond where that is expected; it goes on from a checkpoint when the alignment is known (scheduling only: a
             * sweep that stops too early just continues, one that is not stopped is cut back like any other). */
            J.rowLimit = 0; sd.pausedOn = 0;
            if (rowLimits && reachKnown() && sd.mode <= 1) {
                double bestDist = 1e30; int bi = -1;
                const s64 dj = (s64)aPos1 - (s64)aPos2;
                for (int zz = 0; zz < (int)lanes.size(); zz++) {
                    gx_lane_state& lo = lanes[zz];
                    if (!lo.busy || lo.anchor >= ln.anchor || lo.unsure) continue;     /* (an anchor that may yet be skipped is no reason to stop) */
                    const u32 ip = apos1[lo.anchor];
                    if ((side == 0) != (ip < aPos1) || ip == aPos1) continue;
                    if (llabs(((s64)ip - (s64)apos2[lo.anchor]) - dj) > 4000) continue;
                    const double d = fabs((double)ip - (double)aPos1);
                    if (d < bestDist) { bestDist = d; bi = zz; }
                }
                if (bi >= 0) {
                    gx_side& fs = lanes[bi].s[1 - side];             /* its sweep that comes this way */
                    const double ext = fs.phase == SIDE_DONE ? (double)fs.res.end1 : reachNow();
                    const double every_ = (double)B.ckpt_every();
                    const double lim = std::max(bestDist - ext + 0.02 * ext + 3 * std::max(slackRows, 0.0), 4 * every_);
                    const double already = rec >= 0 ? ((double)rec + 1) * every_ : 0;
                    if (lim < 0.9 * reachNow() && lim > already + 2 * every_) { J.rowLimit = (u32)lim; sd.pausedOn = lanes[bi].anchor; }
                }
            }
            sd.snapshot = G.committed.size();
        }
        (void)m;
        J.al = B.aligns();
        J.tbOnly = tbOnly ? 1 : 0;
        J.token = ++tokenCounter; J.done = 0;
        sd.token = J.token; sd.tbOnly = tbOnly; sd.phase = SIDE_RUNNING;
        batch[sd.mode].push_back((u16)(2 * z + side));
        return 0;
    };

    /* retire the anchors that lie on alignment `ai` (index into G.al; n = the trivial self alignment) */
    bool startDirty = true;                                  /* something happened that may let another anchor start */
    std::vector<u32>* pendp = NULL; bool pendReady = false;   /* (the list is declared further down) */
    /* anchors that wait for an earlier open anchor are parked under it and come back when it is resolved or loses its lane */
    std::unordered_map<u64, std::vector<u32> > waiters; bool pendUnsorted = false;
    auto release_waiters = [&](u64 i) {
        auto it = waiters.find(i);
        if (it == waiters.end()) return;
        if (pendReady) { pendp->insert(pendp->end(), it->second.begin(), it->second.end()); pendUnsorted = true; }
        waiters.erase(it);
    };
    /* take a lane away from its anchor: at once if nothing of it is running, else when its sweeps have stopped */
    auto drop_lane = [&](int z) {
        gx_lane_state& ln = lanes[z];
        laneOf[ln.anchor] = -1; ln.busy = false;
        if (!fin[ln.anchor] && pendReady) { pendp->push_back((u32)ln.anchor); pendUnsorted = true; }   /* back among the waiting */
        release_waiters(ln.anchor);
        for (int side = 0; side < 2; side++) {
            ln.s[side].res.ops.clear();
            if (ln.s[side].phase == SIDE_RUNNING) B.job(z, side)->abort = 1; else ln.s[side].phase = SIDE_IDLE;
        }
        pfWasted++; startDirty = true;
    };
    auto lane_free = [&](int z) { return !lanes[z].busy && lanes[z].s[0].phase != SIDE_RUNNING && lanes[z].s[1].phase != SIDE_RUNNING; };
    auto retire_covered = [&](int ai) {
        galn& x = G.al[ai];
        const int ns = (int)x.segs.size();
        if (ns == 0) return;
        size_t lo = std::lower_bound(byPos.begin(), byPos.end(), x.pos1, [&](u32 a, u32 v) { return apos1[a] < v; }) - byPos.begin();
        for (size_t z = lo; z < byPos.size() && apos1[byPos[z]] <= x.end1; z++) {
            const u32 j = byPos[z];
            if (fin[j] || (int)j == ai) continue;
            const u32 pos1 = apos1[j], pos2 = apos2[j];
            int k = 0, hi2 = ns;                             /* first segment whose e1 >= pos1 */
            while (k < hi2) { int mid = (k + hi2) >> 1; if (x.segs[mid].e1 < pos1) k = mid + 1; else hi2 = mid; }
            if (k == ns) continue;
            const hseg& bp = x.segs[k]; s32 d;
            if (bp.type == SEG_DIAG) d = (s32)(bp.b2 - pos2) + (s32)(pos1 - bp.b1); else d = (s32)(bp.b2 - pos2);
            if (d != 0) continue;
            fin[j] = 1;                                      /* on an alignment committed earlier in the order: skipped for good (:1335) */
            if (laneOf[j] >= 0) { GX_TRACE("[gx %.4f] retire a=%u pos1=%u (on alignment a=%d) while it held a lane\n", now(), j, pos1, ai); drop_lane(laneOf[j]); }
        }
    };
    if (G.obi == (int)n) retire_covered((int)n);

    u64 pairedBases = 0;
    /* commit anchor i from its finished, validated sweeps */
    auto commit_anchor = [&](u64 i, dp_result& rl, dp_result& rr) {
        galn& m = G.al[i];
        G.st.anchorsExtended++;
        /* the reference's counters see exactly the DPs whose results are used (gapped_extend.c:3593,3776) */
        G.st.dpCells += rl.cells + rr.cells; G.st.dpRows += (u64)rl.rows + rr.rows;
        G.st.truncated += (rl.status == DP_TRUNCATED) + (rr.status == DP_TRUNCATED);
        GX_TRACE("[gx %.4f] commit a=%llu pos1=%u rows=%u+%u\n", now(), (unsigned long long)i, m.pos1, rl.rows, rr.rows);
        u32 a1 = m.pos1, a2 = m.pos2;
        u32 start1 = a1 + 1 - rl.end1, start2 = a2 + 1 - rl.end2, stop1 = a1 + rr.end1, stop2 = a2 + rr.end2;
        lzb_editscript* sl = es_new((u32)(rl.ops.size() + rr.ops.size() + 4));
        /* left script: ops in emission order; right script: emitted far-end first, so reversed (:2529-2551) */
        for (size_t k = 0; k < rl.ops.size(); k++) es_add(&sl, rl.ops[k] & 3, rl.ops[k] >> 2);
        for (size_t k = rr.ops.size(); k-- > 0;) es_add(&sl, rr.ops[k] & 3, rr.ops[k] >> 2);
        if (sl->len > 0 && rr.ops.size() > 0) sl->tailOp = rr.ops.back() & 3;
        s32 score = rl.score + rr.score;
        if (sl->len != 0) {
            if ((sl->op[0] & 3) != LZB_OP_SUB) {             /* lop_initial_indels :2589 */
                u32 p1 = start1, p2 = start2, k = 0;
                for (; k < sl->len; k++) { u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2; if (op == LZB_OP_SUB) break; if (op == LZB_OP_INS) p2 += rpt; else p1 += rpt; }
                if (k == sl->len) score = (s32)(-0x7FFFFFFF - 1);
                else { start1 = p1; start2 = p2; sl->len -= k; memmove(sl->op, sl->op + k, (size_t)sl->len * 4); score = rescore(G, start1, start2, sl); }
            }
            if (score != (s32)(-0x7FFFFFFF - 1) && (sl->op[sl->len - 1] & 3) != LZB_OP_SUB) {   /* lop_final_indels :2640 */
                u32 p1 = stop1, p2 = stop2, k = sl->len;
                while (k > 0) { k--; u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2; if (op == LZB_OP_SUB) { k++; break; } if (op == LZB_OP_INS) p2 -= rpt; else p1 -= rpt; }
                if (k == 0) score = (s32)(-0x7FFFFFFF - 1);
                else { stop1 = p1; stop2 = p2; sl->len = k; score = rescore(G, start1, start2, sl); }
            }
        }
        /* format_alignment :5153 */
        u32 beg1 = start1 + 1, end1 = stop1 + 1, beg2 = start2 + 1, end2 = stop2 + 1;
        u32 height = end1 - beg1 + 1, width = end2 - beg2 + 1, k = 0;
        m.segs.clear();
        for (u32 ii = 0, jj = 0; ii < height || jj < width;) {
            u32 si = ii, sj = jj, run = 0;
            while (k < sl->len && (sl->op[k] & 3) == LZB_OP_SUB) { run += sl->op[k] >> 2; k++; }
            ii += run; jj += run;
            add_diag(m, beg1 + si - 1, beg2 + sj - 1, beg1 + ii - 2, beg2 + jj - 2);
            if (ii < height || jj < width) {
                if (k < sl->len) { u32 op = sl->op[k] & 3, rpt = sl->op[k] >> 2; if (op == LZB_OP_INS) jj += rpt; else if (op == LZB_OP_DEL) ii += rpt; k++; }
                else break;
            }
        }
        lzb_alignel* a = (lzb_alignel*)calloc(1, sizeof *a);
        a->script = sl; a->beg1 = beg1; a->beg2 = beg2; a->end1 = end1; a->end2 = end2;
        a->seq1 = h1; a->seq2 = h2; a->s = score; a->hspId = m.hspId;
        m.align = a; m.pos1 = start1; m.pos2 = start2; m.end1 = stop1; m.end2 = stop2;
        fin[i] = 1;
        if (m.segs.empty()) return;
        if (!P->allBounds && a->s < P->scoreThreshold) { free(a->script); free(a); m.align = NULL; m.segs.clear(); return; }
        alignment_neighbours(G, m);
        list_insert(G, (int)i);
        m.devIx = (int)G.committed.size(); G.committed.push_back((int)i);
        retire_covered((int)i);
        if (P->maxPairedBases > 0) {
            /* the limit on aligned columns (gapped_extend.c:1444-1459, count_paired_bases :5695): commits come in the reference's
             * anchor order, so the alignment that takes the sum over the limit is the same one; nothing after it is extended --
             * every anchor still open is closed, the sweeps in flight are told to stop */
            for (const hseg& sg : m.segs) if (sg.type == SEG_DIAG) pairedBases += (u64)sg.e1 + 1 - sg.b1;
            if (pairedBases > P->maxPairedBases) {
                G.st.overlyPaired = 1;
                for (u64 j = 0; j < n; j++) if (!fin[j]) { fin[j] = 1; if (laneOf[j] >= 0) drop_lane(laneOf[j]); }
            }
        }
    };

    /* expected length of a sweep in rows, for deciding what is worth starting: an average over finished sweeps */
    /* How far a sweep will go (scheduling only).  A sweep that is not stopped by anything ends where the traceback
     * runs out (gapped_extend.c:3636-3662): rows ~ tbLen / (traceback bytes per row), and the bytes per row settle
     * within the first thousand rows -- the kernel reports (rows, bytes) with every checkpoint, so the estimate is there
     * a few milliseconds after the first launch and exact as soon as one truncated sweep has finished.  (Where the
     * homology ends earlier the estimate is too long, which only makes a later anchor wait for an earlier one.) */
    const double reachPrior = (double)tbLen / 250.0 + 64;    /* before anything is known: generous */
    double reachTrunc = 0; bool reachExact = false; u32 reachSeen = 0;   /* rows the current estimate was taken over */
    auto reach = [&]() -> double { return reachTrunc > 0 ? reachTrunc : reachPrior; };
    /* how far off the expected reach may be: a finished sweep gives the row itself (sweeps of one call end within a few
     * hundred rows of each other); an estimate is as good as the stretch of rows it was taken over */
    /* (measured on the 50 Mbp pair: the estimate from 4096 rows is 0.25 % off, from 16k rows 0.02 %; finished sweeps differ by +-100 rows) */
    auto slack_frac = [&]() -> double { return reachExact ? 0.001 : reachSeen >= 16384 ? 0.002 : 0.006; };

    /* first row of lane z's sweep `side` in which an alignment committed since the sweep's snapshot shows up (0xFFFFFFFF: none) */
    auto first_touched_row = [&](int z, int side) -> u32 {
        gx_lane_state& ln = lanes[z]; gx_side& sd = ln.s[side];
        const u32 aPos1 = apos1[ln.anchor];
        u32 firstRow = 0xFFFFFFFFu;
        for (size_t k = sd.snapshot; k < G.committed.size(); k++) {
            galn& x = G.al[G.committed[k]];
            if (side == 1) { if (x.pos1 > aPos1) firstRow = std::min(firstRow, x.pos1 - aPos1); }
            else { if (x.end1 < aPos1) firstRow = std::min(firstRow, aPos1 - x.end1); }
        }
        return firstRow;
    };
    /* last checkpoint whose row lies before firstRow: record r holds the state after row (r+1)*every; -1: none */
    auto record_before = [&](gx_side& sd, u32 firstRow) -> int {
        if (sd.ckptCount == 0 || sd.ckptEvery == 0) return -1;
        const u32 full = firstRow == 0xFFFFFFFFu ? sd.ckptCount : (firstRow - 1) / sd.ckptEvery;     /* checkpoints at rows <= firstRow - 1 */
        return (int)std::min<u32>(full, sd.ckptCount) - 1;
    };
    /* validate a DONE side against the alignments committed since its snapshot; returns 0 valid, 1 relaunched, -1 error */
    auto validate_side = [&](int z, int side) -> int {
        gx_lane_state& ln = lanes[z]; gx_side& sd = ln.s[side];
        if (sd.phase != SIDE_DONE || sd.snapshot == G.committed.size()) return 0;
        const u32 firstRow = first_touched_row(z, side);
        if (firstRow > sd.res.rows) { sd.snapshot = G.committed.size(); return 0; }      /* the sweep ended before that row */
        G.st.redone++;
        const int rec = record_before(sd, firstRow);
        if (rec >= 0) pfResumes++; else pfRestarts++;
        GX_TRACE("[gx %.4f] a=%llu side=%d touched at row %u of %u: %s %d\n", now(), (unsigned long long)ln.anchor, side, firstRow, sd.res.rows, rec >= 0 ? "resume from record" : "restart", rec);
        sd.res.ops.clear();
        if (queue_side(z, side, rec, false)) return -1;
        return 1;
    };
    /* a sweep that stopped short of an earlier anchor's expected alignment goes on once that anchor is resolved */
    auto continue_paused = [&](int z, int side) -> int {
        gx_lane_state& ln = lanes[z]; gx_side& sd = ln.s[side];
        const int rec = record_before(sd, first_touched_row(z, side));
        GX_TRACE("[gx %.4f] a=%llu side=%d paused at row %u goes on from record %d\n", now(), (unsigned long long)ln.anchor, side, sd.res.rows, rec);
        pfResumes++;
        return queue_side(z, side, rec, false);
    };

    auto start_anchor = [&](int z, u64 j, bool unsure) -> int {
        gx_lane_state& ln = lanes[z]; galn& y = G.al[j];
        ln.busy = true; ln.anchor = j; ln.left1 = y.left1; ln.right1 = y.right1; ln.unsure = unsure;
        laneOf[j] = z;
        for (int side = 0; side < 2; side++) { ln.s[side].mode = firstMode; ln.s[side].ckptCount = 0; ln.s[side].ckptEvery = 0; ln.s[side].res.ops.clear(); ln.s[side].prog0Rows = ln.s[side].prog0Used = 0; }
        B.job(z, 0)->abort = 0; B.job(z, 1)->abort = 0;
        GX_TRACE("[gx %.4f] start a=%llu pos1=%u lane=%d committed=%zu unsure=%d\n", now(), (unsigned long long)j, y.pos1, z, G.committed.size(), (int)unsure);
        for (int side = 0; side < 2; side++) { ln.s[side].phase = SIDE_IDLE; if (queue_side(z, side, -1, false)) return -1; }
        return 0;
    };

    /* a finished launch of lane z, side: collect it, or send it again with what it asked for */
    auto harvest_side = [&](int z, int side) -> int {
        gx_lane_state& ln = lanes[z]; gx_side& sd = ln.s[side];
        dp_job& J = *B.job(z, side);
        if (!ln.busy) { sd.phase = SIDE_IDLE; startDirty = true; return 0; }      /* the anchor was retired while this ran */
        if (!sd.tbOnly) G.st.dpCellsComputed += J.cells;
        int again = 0;                                        /* 1 fresh rerun, 2 traceback only */
        int ringRec = -1;
        if (J.status == DP_RING) {
            if (sd.mode >= 3) return fail("Y-drop band wider than %u columns; lower --ydrop", B.ring(sd.mode));
            /* the one-warp kernel's 512-column window is outgrown (typically beside an earlier alignment): the four-warp
             * kernel takes over from the last checkpoint; beyond its 1024 columns the shared-memory kernel starts afresh */
            if (sd.mode == 1) { sd.mode = 0; again = 1; ringRec = (int)J.ckptCount - 1; }
            else { sd.mode = sd.mode < 2 ? 2 : sd.mode + 1; again = 1; }
        } else if (J.status == DP_TBROW || J.status == DP_ACT) { if (B.grow(z, side, J.status)) return -1; again = 1; }
        else if (J.opsOverflow) { if (B.grow(z, side, DP_OPS)) return -1; again = 2; }
        if (again) {
            GX_TRACE("[gx %.4f] rerun a=%llu side=%d status=%d overflow=%d mode=%d\n", now(), (unsigned long long)ln.anchor, side, J.status, J.opsOverflow, sd.mode);
            return queue_side(z, side, again == 2 ? -1 : ringRec, again == 2);
        }
        dp_result& r = sd.res;
        r.score = J.score; r.end1 = J.end1; r.end2 = J.end2; r.rows = J.rows; r.status = J.status; r.cells = J.cells;
        const u32* ops = B.ops(z, side);
        r.ops.assign(ops, ops + J.nops);
        sd.ckptCount = J.ckptCount; sd.ckptEvery = sd.mode <= 1 ? B.ckpt_every() : 0;
        sd.phase = J.status == DP_PAUSED ? SIDE_PAUSED : SIDE_DONE;
        if (r.status == DP_TRUNCATED) { reachTrunc = reachExact ? 0.8 * reachTrunc + 0.2 * r.rows : (double)r.rows; if (!reachExact) startDirty = true; reachExact = true; }
        GX_TRACE("[gx %.4f] done a=%llu side=%d rows=%u end1=%u status=%d mode=%d ckpts=%u\n", now(), (unsigned long long)ln.anchor, side, r.rows, r.end1, r.status, sd.mode, sd.ckptCount);
        return 0;
    };

    /* ---- the anchor loop, gapped_extend.c:1300-1470 ---- */
    std::vector<int> blocker(n, -1);                         /* the earlier unresolved anchor this one is probably covered by */
    std::vector<u32> pend(n);                                /* anchors neither resolved nor started, best first (compacted as it is walked) */
    for (u64 i = 0; i < n; i++) pend[i] = (u32)i;
    double lastSlackFrac = -1;
    pendp = &pend; pendReady = true;
    u64 hd = 0; double lastEventAt = -1;
    while (true) {
        while (hd < n && fin[hd]) hd++;
        if (hd >= n) break;
        bool progressed = false;
        if (lastEventAt < 0) lastEventAt = now();
        /* 1. collect finished launches */
        double t0 = prof ? now() : 0;
        for (int z = 0; z < have; z++) for (int side = 0; side < 2; side++) {
            gx_side& sd = lanes[z].s[side];
            if (sd.phase != SIDE_RUNNING) continue;
            dp_job& J = *B.job(z, side);
            if (J.done != sd.token) continue;
            std::atomic_thread_fence(std::memory_order_acquire);
            if (harvest_side(z, side)) return -1;
            progressed = true;
        }
        /* 2. check finished sweeps against what has been committed since; resume the ones that were reached */
        for (int z = 0; z < have; z++) {
            gx_lane_state& ln = lanes[z];
            if (!ln.busy) continue;
            const bool anyRunning = ln.s[0].phase == SIDE_RUNNING || ln.s[1].phase == SIDE_RUNNING;
            bool stale = false, paused = false;
            for (int side = 0; side < 2; side++) {
                if (ln.s[side].phase == SIDE_DONE && ln.s[side].snapshot != G.committed.size()) stale = true;
                if (ln.s[side].phase == SIDE_PAUSED && (fin[ln.s[side].pausedOn] || laneOf[ln.s[side].pausedOn] < 0)) paused = true;
            }
            if (!stale && !paused) continue;
            /* an alignment across the anchor's row may be a new neighbour (or cover the anchor: retire_covered saw to that) */
            galn& y = G.al[ln.anchor];
            bool crossed = false;
            const size_t from = std::min(ln.s[0].phase != SIDE_IDLE ? ln.s[0].snapshot : G.committed.size(),
                                         ln.s[1].phase != SIDE_IDLE ? ln.s[1].snapshot : G.committed.size());
            for (size_t k = from; k < G.committed.size() && !crossed; k++) { galn& x = G.al[G.committed[k]]; if (x.pos1 <= apos1[ln.anchor] && x.end1 >= apos1[ln.anchor]) crossed = true; }
            if (crossed) {
                int coverer = -1;
                if (!anchor_neighbours(G, y, &coverer)) return fail("internal error: anchor %llu lies on alignment %d but was not retired", (unsigned long long)ln.anchor, coverer);
                if (y.left1.al != ln.left1.al || y.left1.sg != ln.left1.sg || y.right1.al != ln.right1.al || y.right1.sg != ln.right1.sg) {
                    if (anyRunning) continue;                 /* both sweeps restart once the running one is back */
                    GX_TRACE("[gx %.4f] a=%llu new neighbours: restart\n", now(), (unsigned long long)ln.anchor);
                    G.st.redone += 2; pfRestarts += 2;
                    ln.left1 = y.left1; ln.right1 = y.right1;
                    for (int side = 0; side < 2; side++) { ln.s[side].res.ops.clear(); if (queue_side(z, side, -1, false)) return -1; }
                    progressed = true;
                    continue;
                }
            }
            for (int side = 0; side < 2; side++) {              /* (a sweep that is still running is looked at when it is back) */
                if (ln.s[side].phase == SIDE_PAUSED) {
                    if (fin[ln.s[side].pausedOn] || laneOf[ln.s[side].pausedOn] < 0) { if (continue_paused(z, side)) return -1; progressed = true; }
                    continue;
                }
                const int rc = validate_side(z, side); if (rc < 0) return -1; if (rc) progressed = true;
            }
        }
        if (prof) pfValidate += now() - t0;
        /* 3. commit the head anchor while it is ready */
        t0 = prof ? now() : 0;
        {
            const int z = laneOf[hd];
            if (z >= 0) {
                gx_lane_state& ln = lanes[z];
                if (ln.s[0].phase == SIDE_DONE && ln.s[1].phase == SIDE_DONE && ln.s[0].snapshot == G.committed.size() && ln.s[1].snapshot == G.committed.size()) {
                    const u64 i = hd;
                    commit_anchor(i, ln.s[0].res, ln.s[1].res);
                    release_waiters(i);
                    ln.busy = false; ln.s[0].phase = ln.s[1].phase = SIDE_IDLE; ln.s[0].res.ops.clear(); ln.s[1].res.ops.clear(); laneOf[i] = -1;
                    startDirty = true;
                    if (prof) pfCommit += now() - t0;
                    if (flush()) return -1;
                    continue;                                 /* later lanes must be re-validated before anything else */
                }
            }
        }
        if (prof) pfCommit += now() - t0;
        /* 4. start anchors, best first, while lanes are free.  The head anchor always gets one. */
        t0 = prof ? now() : 0;
        /* no sweep has finished yet to say how far sweeps go: the running ones report their progress.  Two reports of one
         * sweep give the traceback bytes a row takes once the band has settled, hence the row where the traceback will
         * run out; the longer the sweep has run, the better the estimate. */
        if (!reachExact) {
            for (int z = 0; z < have; z++) for (int side = 0; side < 2; side++) {
                gx_side& sd = lanes[z].s[side];
                if (!lanes[z].busy || sd.phase != SIDE_RUNNING) continue;
                dp_job& J = *B.job(z, side);
                const u32 pr = J.progRows, pu = J.progUsed;
                if (J.resume >= 0 || pr < 2048 || pu == 0) continue;
                if (sd.prog0Rows == 0) { sd.prog0Rows = pr; sd.prog0Used = pu; continue; }
                if (pr < sd.prog0Rows + 4096 || pu <= sd.prog0Used || pr - sd.prog0Rows <= reachSeen) continue;
                const double perRow = (double)(pu - sd.prog0Used) / (double)(pr - sd.prog0Rows);
                const double before = slack_frac();
                reachTrunc = pr + ((double)tbLen - pu - 2.0 * perRow) / perRow;
                reachSeen = pr - sd.prog0Rows;
                if (slack_frac() != before || (reachSeen >> 11) > ((reachSeen - 256) >> 11)) GX_TRACE("[gx %.4f] estimate a=%llu: sweeps will run %.0f rows (from %u rows of a=%llu side %d)\n", now(), (unsigned long long)lanes[z].anchor, reachTrunc, reachSeen, (unsigned long long)lanes[z].anchor, side);
            }
            /* the first estimate, and every sharper one, lets more anchors be judged */
            if (reachTrunc > 0 && slack_frac() != lastSlackFrac) startDirty = true;
        }
        if (startDirty) {
            startDirty = false; pfPasses++;
            int freeLanes = 0;
            for (int z = 0; z < have; z++) if (lane_free(z)) freeLanes++;
            auto more_lanes = [&]() -> int {                    /* made as they are needed */
                if (have >= W) return 0;
                const int got = B.lanes(std::min(W, have + 32));
                if (got < 0) return -1;
                if (got == have) W = have;                      /* the device has no room for more */
                freeLanes += got - have; have = got;
                return 0;
            };
            if (laneOf[hd] < 0 && freeLanes == 0 && more_lanes()) return -1;
            if (laneOf[hd] < 0 && freeLanes == 0) {
                /* every lane is held by a later anchor: take back the latest one that is not running */
                int victim = -1;
                for (int z = 0; z < have; z++) {
                    gx_lane_state& ln = lanes[z];
                    if (!ln.busy || ln.s[0].phase == SIDE_RUNNING || ln.s[1].phase == SIDE_RUNNING) continue;
                    if (victim < 0 || ln.anchor > lanes[victim].anchor) victim = z;
                }
                if (victim >= 0) {
                    GX_TRACE("[gx %.4f] lane %d taken back from a=%llu for the head anchor %llu\n", now(), victim, (unsigned long long)lanes[victim].anchor, (unsigned long long)hd);
                    drop_lane(victim); startDirty = false;
                    freeLanes = 1;
                } else startDirty = true;                     /* look again when a sweep has finished */
            }
            const bool calibrated = reachTrunc > 0;
            const double rr = reach();
            reachShared = calibrated ? rr : 0;
            const double slackFrac = slack_frac();
            lastSlackFrac = calibrated ? slackFrac : -1;
            std::vector<std::pair<u32, int> > lanePos;
            for (int z = 0; z < have; z++) if (lanes[z].busy) lanePos.push_back(std::make_pair(apos1[lanes[z].anchor], z));
            std::sort(lanePos.begin(), lanePos.end());
            if (pendUnsorted) { std::sort(pend.begin(), pend.end()); pend.erase(std::unique(pend.begin(), pend.end()), pend.end()); pendUnsorted = false; }
            size_t keep = 0, q = 0;
            int unsureRunning = 0;
            for (int z = 0; z < have; z++) if (lanes[z].busy && lanes[z].unsure) unsureRunning++;
            for (; q < pend.size() && (freeLanes > 0 || have < W); q++) {
                const u64 j = pend[q];
                if (fin[j] || laneOf[j] >= 0) continue;          /* resolved or started since: leaves the list */
                pend[keep++] = (u32)j;
                bool unsure = false;
                if (j != hd && slackRows >= 0) {
                    /* Will an earlier anchor that is still open come to cover this one?  (Scheduling only.)
                     *   well inside its expected reach: yes -- wait for its commit (the reference skips such anchors, :1330);
                     *   near the edge of its reach: nobody knows yet -- it starts (its sweep towards the earlier anchor stops short
                     *     of that anchor's expected alignment, see queue_side), but as it may yet be skipped it holds nobody back,
                     *     and only so many lanes go to such anchors;
                     *   before any sweep has said how far sweeps go, only anchors far from every started one start. */
                    blocker[j] = -1;
                    const s64 dj = (s64)apos1[j] - (s64)apos2[j];
                    bool tooEarly = false;
                    const double zone = calibrated ? 1.06 * rr + slackRows + 16 : 2.5 * reachPrior;
                    const u32 pj = apos1[j];
                    const size_t mid = std::lower_bound(lanePos.begin(), lanePos.end(), std::make_pair(pj, -1)) - lanePos.begin();
                    for (int dir = 0; dir < 2 && blocker[j] < 0; dir++)
                    for (size_t w = dir ? mid : mid - 1; w < lanePos.size() && blocker[j] < 0; dir ? w++ : w--) {
                        if (fabs((double)lanePos[w].first - (double)pj) > zone) break;
                        const int z = lanePos[w].second;
                        gx_lane_state& ln = lanes[z];
                        if (!ln.busy || ln.anchor >= j) continue;
                        const u64 i = ln.anchor;
                        const s64 di = (s64)apos1[i] - (s64)apos2[i];
                        if (llabs(di - dj) > 4000) continue;
                        const int side = apos1[j] < apos1[i] ? 0 : 1;
                        const double dist = fabs((double)apos1[j] - (double)apos1[i]);
                        if (!calibrated) { tooEarly = true; continue; }
                        const bool known = ln.s[side].phase == SIDE_DONE;
                        const double ext = known ? (double)ln.s[side].res.end1 : rr;
                        const double slack = (known ? 0.001 : slackFrac) * ext + slackRows;
                        if (dist <= ext - slack) { if (ln.unsure) unsure = true; else blocker[j] = (int)i; }
                        else if (dist <= ext + slack) unsure = true;
                        if (blocker[j] >= 0 && j < 6000) GX_TRACE("[gx %.4f] wait a=%llu pos1=%u for a=%llu (dist %.0f of %.0f, side %d %s)\n", now(), (unsigned long long)j, apos1[j], (unsigned long long)i, dist, ext, side, known ? "done" : "running");
                    }
                    if (blocker[j] >= 0) { waiters[(u64)blocker[j]].push_back((u32)j); keep--; continue; }     /* parked */
                    if (tooEarly) continue;
                    if (unsure && unsureRunning >= std::max(4, W / 4)) continue;
                }
                galn& y = G.al[j];
                int coverer = -1;
                if (!anchor_neighbours(G, y, &coverer)) { fin[j] = 1; progressed = true; continue; }     /* (retire_covered normally got there first) */
                if (freeLanes == 0) { if (more_lanes()) return -1; if (freeLanes == 0) break; }
                int fl = -1;
                for (int z = 0; z < have; z++) if (lane_free(z)) { fl = z; break; }
                if (start_anchor(fl, j, unsure)) return -1;
                if (unsure) unsureRunning++;
                lanePos.insert(std::lower_bound(lanePos.begin(), lanePos.end(), std::make_pair(apos1[j], fl)), std::make_pair(apos1[j], fl));
                freeLanes--; progressed = true; keep--;
                if (j != hd) G.st.speculated++;
            }
            for (; q < pend.size(); q++) pend[keep++] = pend[q];
            pend.resize(keep);
        }
        if (prof) pfStart += now() - t0;
        if (flush()) return -1;
        if (progressed) { lastEventAt = now(); continue; }
        /* 5. nothing to decide: wait for the device */
        bool any = false;
        for (int z = 0; z < have && !any; z++) if (lanes[z].s[0].phase == SIDE_RUNNING || lanes[z].s[1].phase == SIDE_RUNNING) any = true;
        if (!any) return fail("internal error: gapped scheduler stalled at anchor %llu", (unsigned long long)hd);
        t0 = prof ? now() : 0;
        if (!B.poll()) return fail("Y-drop kernel failed: %s", B.error());
        if (prof) { pfPoll += now() - t0; pfPolls++; }
        if (now() - lastEventAt > 180.0) return fail("the gapped stage made no progress for 180 s (anchor %llu)", (unsigned long long)hd);
    }
    /* sweeps of retired anchors that are still running: ask them to stop and wait (their buffers are reused by the next call) */
    for (int z = 0; z < have; z++) for (int side = 0; side < 2; side++) if (lanes[z].s[side].phase == SIDE_RUNNING) B.job(z, side)->abort = 1;
    for (int z = 0; z < have; z++) for (int side = 0; side < 2; side++) {
        while (lanes[z].s[side].phase == SIDE_RUNNING) {
            if (B.job(z, side)->done == lanes[z].s[side].token) { lanes[z].s[side].phase = SIDE_IDLE; break; }
            if (!B.poll()) return fail("Y-drop kernel failed: %s", B.error());
        }
    }
    if (trace) fwrite(traceBuf.data(), 1, traceBuf.size(), stderr);
    if (prof)
        fprintf(stderr, "[gx profile] lanes=%d wall=%.3f validate_s=%.3f commit_s=%.3f start_s=%.3f poll_s=%.3f polls=%llu launches=%llu jobs=%llu resumes=%llu restarts=%llu "
                        "retired_while_held=%llu extended=%llu redone=%llu rows=%llu committed=%zu reach=%.0f seen=%u passes=%llu\n",
                have, now(), pfValidate, pfCommit, pfStart, pfPoll, (unsigned long long)pfPolls, (unsigned long long)pfLaunches, (unsigned long long)pfJobs,
                (unsigned long long)pfResumes, (unsigned long long)pfRestarts, (unsigned long long)pfWasted, (unsigned long long)G.st.anchorsExtended,
                (unsigned long long)G.st.redone, (unsigned long long)G.st.dpRows, G.committed.size(), reach(), reachSeen, (unsigned long long)pfPasses);
    lzb_alignel* head = NULL, *last = NULL;
    for (int o = G.obi; o >= 0; o = G.al[o].next) {
        galn& m = G.al[o];
        bool drop = m.align->s < P->scoreThreshold || (P->inhibitTrivial && m.align->isTrivial);
        if (G.st.overlyPaired && !P->overlyPairedKeep) drop = true;       /* discard_alignments :1580 */
        if (drop) { free(m.align->script); free(m.align); }
        else { if (!head) head = last = m.align; else { last->next = m.align; last = m.align; } }
    }
    *list = head;
    G.st.seconds = now();
    if (stats) *stats = G.st;
    return 0;
}

#endif
