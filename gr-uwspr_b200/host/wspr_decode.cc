// Host side of the path that stays on the CPU: symbol de-interleaving and the K=32 r=1/2
// sequential (Fano) decoder that sync_and_demodulate hands the soft symbols to.
//
// Reference behaviour being matched (file:line under the reference tree):
//   lib/sync_and_demodulate_impl.cc:265-282   deinterleave (8-bit bit-reversal order)
//   lib/Fano.cc:36-45                         integer branch-metric table (wspr_tables.h)
//   lib/Fano.cc:110-252                       decoder; metric, cycle and progress counters
//                                             are reproduced so callers can log them
//   lib/sync_and_demodulate_impl.cc:457-490   peak-up / decode loop of one candidate
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "uwspr_b200.h"
#include "wspr_tables.h"

namespace {

const uint32_t kPoly1 = 0xf2d05351u;  // Layland-Lushbaugh polynomials, lib/Fano.cc:54-55
const uint32_t kPoly2 = 0xe4613c47u;

inline uint32_t parity(uint32_t v) { return (uint32_t)__builtin_parity(v); }

// channel symbol pair of an encoder state: POLY1 parity in bit 1, POLY2 parity in bit 0
inline uint32_t branch_pair(uint64_t state)
{
    const uint32_t s = (uint32_t)state;
    return (parity(s & kPoly1) << 1) | parity(s & kPoly2);
}

struct Node {
    uint64_t state;   // encoder state of the next node
    int64_t gamma;    // cumulative metric up to this node
    int metric[4];    // branch metrics for the four possible symbol pairs
    int sorted[2];    // metrics of the two branches, best first
    int branch;       // which of the two is being explored
};

// writes order[p] = source index of output p
void bit_reversal_order(uint8_t order[UWSPR_B200_NSYM])
{
    int p = 0;
    for (int i = 0; p < UWSPR_B200_NSYM && i < 256; i++) {
        int j = 0;
        for (int b = 0; b < 8; b++)
            if (i & (1 << b)) j |= 1 << (7 - b);
        if (j < UWSPR_B200_NSYM) order[p++] = (uint8_t)j;
    }
}

inline void sort_branches(Node &n, uint32_t pair)
{
    const int m0 = n.metric[pair], m1 = n.metric[3 ^ pair];
    if (m0 > m1) {
        n.sorted[0] = m0;
        n.sorted[1] = m1;
    } else {
        n.sorted[0] = m1;
        n.sorted[1] = m0;
        n.state++;  // the 1-branch is the better one
    }
}

}  // namespace

extern "C" void uwspr_b200_deinterleave(uint8_t *sym)
{
    uint8_t order[UWSPR_B200_NSYM], tmp[UWSPR_B200_NSYM];
    bit_reversal_order(order);
    for (int p = 0; p < UWSPR_B200_NSYM; p++) tmp[p] = sym[order[p]];
    memcpy(sym, tmp, UWSPR_B200_NSYM);
}

extern "C" int uwspr_b200_fano(uint32_t *metric, uint32_t *cycles, uint32_t *maxnp, uint8_t *data,
                               const uint8_t *symbols, uint32_t nbits, int delta, uint32_t maxcycles)
{
    std::vector<Node> nodes(nbits + 1);
    memset(nodes.data(), 0, sizeof(Node) * (nbits + 1));
    const int last = (int)nbits - 1, tail = (int)nbits - 31;
    for (int k = 0; k <= last; k++) {
        const int a = symbols[2 * k], b = symbols[2 * k + 1];
        nodes[k].metric[0] = WSPR_METTAB[0][a] + WSPR_METTAB[0][b];
        nodes[k].metric[1] = WSPR_METTAB[0][a] + WSPR_METTAB[1][b];
        nodes[k].metric[2] = WSPR_METTAB[1][a] + WSPR_METTAB[0][b];
        nodes[k].metric[3] = WSPR_METTAB[1][a] + WSPR_METTAB[1][b];
    }
    int np = 0;
    uint32_t deepest = 0;
    sort_branches(nodes[0], branch_pair(nodes[0].state));
    const uint64_t limit = (uint64_t)maxcycles * nbits;
    int64_t threshold = 0;
    uint64_t i;
    bool done = false;
    for (i = 1; i <= limit; i++) {
        if (np > (int)deepest) deepest = (uint32_t)np;
        Node &cur = nodes[np];
        const int64_t ahead = cur.gamma + cur.sorted[cur.branch];
        if (ahead >= threshold) {
            // the node is acceptable; on a first visit tighten the threshold
            if (cur.gamma < threshold + delta)
                while (ahead >= threshold + delta) threshold += delta;
            nodes[np + 1].gamma = ahead;
            nodes[np + 1].state = cur.state << 1;
            if (++np == last + 1) {
                done = true;
                break;
            }
            Node &nx = nodes[np];
            const uint32_t pair = branch_pair(nx.state);
            if (np >= tail)
                nx.sorted[0] = nx.metric[pair];  // the tail is all zeros: no 1-branch
            else
                sort_branches(nx, pair);
            nx.branch = 0;
            continue;
        }
        // threshold violated: back up until a node offers an untried second branch
        for (;;) {
            if (np == 0 || nodes[np - 1].gamma < threshold) {
                threshold -= delta;  // cannot back up: relax and retry the best branch
                if (nodes[np].branch != 0) {
                    nodes[np].branch = 0;
                    nodes[np].state ^= 1;
                }
                break;
            }
            --np;
            if (np < tail && nodes[np].branch != 1) {
                nodes[np].branch++;
                nodes[np].state ^= 1;
                break;
            }
        }
    }
    (void)done;
    *metric = (uint32_t)nodes[np].gamma;
    *maxnp = deepest;
    const uint32_t nbytes = nbits >> 3;
    for (uint32_t b = 0; b < nbytes; b++) data[b] = (uint8_t)nodes[7 + 8 * b].state;
    *cycles = (uint32_t)(i + 1);
    return (i >= limit) ? -1 : 0;
}

extern "C" int uwspr_b200_decode_candidate(const uwspr_b200_refined_t *refined, const uwspr_b200_jiggle_t *jig,
                                           const uint8_t *soft, int jig_count, int8_t *message7,
                                           int32_t *idt_used, uint32_t *fano_cycles)
{
    const uint32_t maxcycles = 10000;  // lib/sync_and_demodulate_impl.cc:329
    const int delta = 60;              // :335
    if (idt_used) *idt_used = -1;
    if (fano_cycles) *fano_cycles = 0;
    if (!refined->worth_a_try) return 0;
    for (int t = 0; t < jig_count; t++) {
        if (!jig[t].gate) continue;    // :475
        uint8_t sym[UWSPR_B200_NSYM], data[11];
        memcpy(sym, soft + (size_t)t * UWSPR_B200_NSYM, UWSPR_B200_NSYM);
        memset(data, 0, sizeof(data));
        uwspr_b200_deinterleave(sym);  // :476
        uint32_t metric, cycles, maxnp;
        const int r = uwspr_b200_fano(&metric, &cycles, &maxnp, data, sym, 81, delta, maxcycles);
        if (fano_cycles) *fano_cycles += cycles;
        if (r == 0) {
            for (int b = 0; b < 7; b++) message7[b] = (int8_t)data[b];  // :484-490
            if (idt_used) *idt_used = t;
            return 1;
        }
    }
    return 0;
}

// Batch form of the loop above.  Work is handed out in blocks of 16 candidates from an
// atomic cursor because decode times are very uneven (about 80 decoder cycles for a clean
// frame, 17 x 10000 for a gated candidate that never decodes).
extern "C" int uwspr_b200_decode_batch(const uwspr_b200_refined_t *refined, const uwspr_b200_jiggle_t *jig,
                                       const uint8_t *soft, int64_t ncand, int jig_count, int nthreads,
                                       uint8_t *decoded, int8_t *messages7, int32_t *idt_used,
                                       uint32_t *fano_cycles)
{
    if (ncand < 0 || jig_count < 0 || jig_count > UWSPR_B200_NJIG) return -UWSPR_B200_E_PARAM;
    if (ncand == 0) return 0;
    if (!refined || !jig || !soft || !decoded || !messages7) return -UWSPR_B200_E_PARAM;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    const int64_t block = 16;
    const int64_t nblocks = (ncand + block - 1) / block;
    if (nthreads > nblocks) nthreads = (int)nblocks;

    std::atomic<int64_t> cursor(0);
    std::atomic<int> total(0);
    auto worker = [&]() {
        int mine = 0;
        for (;;) {
            const int64_t b = cursor.fetch_add(1, std::memory_order_relaxed);
            if (b >= nblocks) break;
            const int64_t end = (b + 1) * block < ncand ? (b + 1) * block : ncand;
            for (int64_t g = b * block; g < end; g++) {
                int32_t idt = -1;
                uint32_t cyc = 0;
                int8_t *msg = messages7 + g * 7;
                memset(msg, 0, 7);
                const int ok = uwspr_b200_decode_candidate(refined + g, jig + g * jig_count,
                                                           soft + (size_t)g * jig_count * UWSPR_B200_NSYM,
                                                           jig_count, msg, &idt, &cyc);
                decoded[g] = (uint8_t)ok;
                if (idt_used) idt_used[g] = idt;
                if (fano_cycles) fano_cycles[g] = cyc;
                mine += ok;
            }
        }
        total.fetch_add(mine, std::memory_order_relaxed);
    };
    if (nthreads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        pool.reserve((size_t)nthreads - 1);
        for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
    }
    return total.load();
}
