// bp_plan.cpp -- host-side planning for the kernels (no CUDA in this file).
//
//   compute_priors        the channel priors log((1-p)/p), computed with the host libm exactly as the reference
//                         does (src_cpp/bp.hpp:150-151)
//   build_serial_batches  levelised serial schedule for the streaming kernels (bp_stream.cuh)
//   build_smem_plan       shared-memory tables of the thread-group kernels (bp_smem.cuh, bp_smem_serial.cuh):
//                         conflict-free message placement by edge colouring, serial levels
#include <algorithm>
#include <cmath>
#include <cstring>

#include "bp_decoder.h"
#include "ref_libm.h"

namespace bpb {

// The product-sum kernels evaluate tanh / log with ref_libm.h, a restatement of glibc 2.39's x86-64 FMA variants, so
// that their messages are the doubles the reference computes on such a host.  On a host whose libm is a different
// one (another glibc, no FMA) the reference itself computes different last bits; this check says so instead of letting
// "bit-exact" silently become "within 1e-5": it evaluates the restatement on the host over a deterministic sample
// (near 1, tiny, large, the tanh -> atanh chain of bp.hpp:208-217) and counts disagreements with the live libm.
int libm_selfcheck(int samples) {
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return s;
    };
    auto uni = [&](double lo, double hi) { return lo + (hi - lo) * ((double) (rnd() >> 11) * 0x1p-53); };
    auto same = [](double a, double b) { return std::memcmp(&a, &b, 8) == 0 || (a != a && b != b); };
    int bad = 0;
    for (int i = 0; i < samples; i++) {
        const double xs[5] = {uni(0.0, 4.0), std::exp(uni(-700.0, 700.0)), uni(0.9, 1.1), uni(-45.0, 45.0),
                              uni(-2.2, 2.2)};
        for (double x: xs) {
            if (!same(rl_log(x), std::log(x))) bad++;
            if (!same(rl_tanh(x), std::tanh(x))) bad++;
        }
        const double b1 = uni(-40, 40), b2 = uni(-6, 6);
        const double c1 = std::tanh(b1 / 2) * std::tanh(b2 / 2), c2 = rl_tanh(b1 / 2) * rl_tanh(b2 / 2);
        if (!same(std::log((1 + c1) / (1 - c1)), rl_log((1 + c2) / (1 - c2)))) bad++;
    }
    return bad;
}

void compute_priors(bpb_decoder *h) {
    const bpb::HostGraph &g = h->g;
    // priors on the host, the reference's expression (bp.hpp:150-151)
    h->prior.resize((size_t) g.n);
    h->uniform_prior = true;
    for (int j = 0; j < g.n; j++) {
        h->prior[(size_t) j] = std::log((1 - h->channel[(size_t) j]) / h->channel[(size_t) j]);
        if (std::memcmp(&h->prior[(size_t) j], &h->prior[0], sizeof(double)) != 0) h->uniform_prior = false;
    }
}

// Levelise the serial schedule (see the serial branch of bp_stream.cuh): level(q) = 1 + max level of the earlier
// schedule positions whose bit shares a check with this one; stable-sort by level; cut each level into batches of
// `sb` bits, padding the last batch of a level with 0xffffffff.
std::vector<uint32_t> build_serial_batches(const bpb::HostGraph &g, const std::vector<uint32_t> &order, int sb) {
    std::vector<int> last_level((size_t) g.m, 0), level(order.size(), 0);
    int max_level = 0;
    for (size_t q = 0; q < order.size(); q++) {
        const uint32_t j = order[q];
        int lv = 0;
        for (uint32_t e = g.col_ptr[j]; e < g.col_ptr[j + 1]; e++) lv = std::max(lv, last_level[g.row_idx[e]]);
        lv += 1;
        for (uint32_t e = g.col_ptr[j]; e < g.col_ptr[j + 1]; e++) last_level[g.row_idx[e]] = lv;
        level[q] = lv;
        max_level = std::max(max_level, lv);
    }
    std::vector<std::vector<uint32_t>> by_level((size_t) max_level + 1);
    for (size_t q = 0; q < order.size(); q++) by_level[(size_t) level[q]].push_back(order[q]);
    std::vector<uint32_t> out;
    for (const auto &bits: by_level) {
        for (uint32_t j: bits) out.push_back(j);
        while (out.size() % (size_t) sb) out.push_back(0xffffffffu);
    }
    return out;
}


static inline uint32_t align_up(uint32_t x, uint32_t q) { return (x + q - 1) / q * q; }

// Conflict-free message placement.  Every message is written by one thread and read by another: row threads touch
// it in the check pass (L consecutive rows, same slot k, form one shared-memory phase), column threads in the bit
// pass (L consecutive columns, same slot).  L = 16 for 8-byte accesses (served per half-warp, conflict-free iff the
// 16 lanes hit 16 different bank pairs), L = 8 for 16-byte accesses (served per quarter-warp, 8 different bank
// quads).  Take the bipartite multigraph whose left nodes are the (row group, slot) cells, right nodes the (column
// group, slot) cells and whose edges are the nonzeros of H: every node has degree <= L, so by Koenig's theorem its
// edges can be coloured with L colours such that no two edges at a node share a colour.  Colour = bank pair / quad:
// both passes become conflict-free.  (Alternating-path edge colouring.)  Returns the length of the message array in
// slots (L * largest colour class); slot_of_edge[CSR edge id] = colour + L * rank inside the colour class.
int place_messages(const HostGraph &g, int L, std::vector<uint32_t> &slot_of_edge) {
    const int DCm = g.max_row_degree, DVm = g.max_col_degree;
    const int n_rc = ((g.m + L - 1) / L) * DCm, n_cc = ((g.n + L - 1) / L) * DVm;
    std::vector<int> edge_rc((size_t) g.nnz), edge_cc((size_t) g.nnz), colour((size_t) g.nnz, -1);
    std::vector<int> at_r((size_t) n_rc * L, -1), at_c((size_t) n_cc * L, -1);  // edge using colour q at the node
    for (int i = 0; i < g.m; i++)
        for (uint32_t q = g.row_ptr[(size_t) i]; q < g.row_ptr[(size_t) i + 1]; q++)
            edge_rc[q] = (i / L) * DCm + (int) (q - g.row_ptr[(size_t) i]);
    for (int j = 0; j < g.n; j++)
        for (uint32_t q = g.col_ptr[(size_t) j]; q < g.col_ptr[(size_t) j + 1]; q++)
            edge_cc[g.csc2csr[q]] = (j / L) * DVm + (int) (q - g.col_ptr[(size_t) j]);
    for (int e = 0; e < g.nnz; e++) {
        const int u = edge_rc[(size_t) e], v = edge_cc[(size_t) e];
        int a = 0, b = 0;
        while (at_r[(size_t) u * L + a] >= 0) a++;  // free at u (exists: degree <= L and e itself uncoloured)
        while (at_c[(size_t) v * L + b] >= 0) b++;  // free at v
        if (a != b) {
            // walk the a/b alternating path that starts at v with colour a and swap a <-> b along it; it cannot
            // reach u (bipartite, a is free at u), so afterwards a is free at both ends
            std::vector<int> path;
            int node = v, want = a;
            bool on_col_side = true;
            for (;;) {
                const int f = on_col_side ? at_c[(size_t) node * L + want] : at_r[(size_t) node * L + want];
                if (f < 0) break;
                path.push_back(f);
                node = on_col_side ? edge_rc[(size_t) f] : edge_cc[(size_t) f];
                on_col_side = !on_col_side;
                want = (want == a) ? b : a;
            }
            for (int f: path) {
                at_r[(size_t) edge_rc[(size_t) f] * L + colour[(size_t) f]] = -1;
                at_c[(size_t) edge_cc[(size_t) f] * L + colour[(size_t) f]] = -1;
            }
            for (int f: path) {
                colour[(size_t) f] = (colour[(size_t) f] == a) ? b : a;
                at_r[(size_t) edge_rc[(size_t) f] * L + colour[(size_t) f]] = f;
                at_c[(size_t) edge_cc[(size_t) f] * L + colour[(size_t) f]] = f;
            }
        }
        colour[(size_t) e] = a;
        at_r[(size_t) u * L + a] = e;
        at_c[(size_t) v * L + a] = e;
    }
    std::vector<int> per_colour((size_t) L, 0);
    slot_of_edge.assign((size_t) g.nnz, 0u);
    for (int e = 0; e < g.nnz; e++)
        slot_of_edge[(size_t) e] = (uint32_t) (colour[(size_t) e] + L * per_colour[(size_t) colour[(size_t) e]]++);
    int longest = 0;
    for (int q = 0; q < L; q++) longest = std::max(longest, per_colour[(size_t) q]);
    return L * longest;
}


void build_smem_plan(bpb_decoder *h) {
    const bpb::HostGraph &g = h->g;
    bpb::SmemPlan &pl = h->smem_plan;
    pl = bpb::SmemPlan();
    const int DCm = g.max_row_degree, DVm = g.max_col_degree;
    const int M = (int) align_up((uint32_t) g.m, 16), N = (int) align_up((uint32_t) g.n, 32);
    if (DCm > 32 || DVm > 16) {
        pl.why = "row degree > 32 or column degree > 16";
        return;
    }
    if (DCm < 1 || (int64_t) g.nnz + 16 * 17 > 65535 || g.n > 65535 || g.m > 65535) {
        pl.why = "message positions do not fit 16-bit indices";
        return;
    }
    pl.M = M;
    pl.N = N;
    uint32_t off = 0;
    pl.off_row_deg = off;
    off += (uint32_t) M;
    pl.off_col_deg = off;
    off += (uint32_t) N;
    off = align_up(off, 4);
    // 16-bit entries, two slots (2q, 2q+1) of the same row / column packed into one 32-bit word: tab[q*stride + x]
    const int DCp = (DCm + 1) / 2, DVp = (DVm + 1) / 2;
    pl.off_col_row = off;
    off += 4u * (uint32_t) (DVp * N);
    pl.off_row_pos = off;
    off += 4u * (uint32_t) (DCp * M);
    pl.off_col_pos = off;
    off += 4u * (uint32_t) (DVp * N);
    // serial schedule: slot of every edge inside its row, and the levelised schedule
    pl.serial = (h->schedule == BPB_SERIAL);
    if (h->serial_order.empty()) {
        h->serial_order.resize((size_t) g.n);
        for (int j = 0; j < g.n; j++) h->serial_order[(size_t) j] = (uint32_t) j;
    }
    std::vector<uint16_t> lev_ptr, lev_bits;
    if (pl.serial) {
        if (h->serial_order.size() > 65535) {
            pl.why = "serial schedule longer than 65535 entries";
            return;
        }
        std::vector<int> last_level((size_t) g.m, 0), level(h->serial_order.size(), 0);
        int max_level = 0;
        for (size_t q = 0; q < h->serial_order.size(); q++) {
            const uint32_t j = h->serial_order[q];
            int lv = 0;
            for (uint32_t e = g.col_ptr[j]; e < g.col_ptr[j + 1]; e++) lv = std::max(lv, last_level[g.row_idx[e]]);
            lv += 1;
            for (uint32_t e = g.col_ptr[j]; e < g.col_ptr[j + 1]; e++) last_level[g.row_idx[e]] = lv;
            level[q] = lv;
            max_level = std::max(max_level, lv);
        }
        lev_ptr.assign((size_t) max_level + 1, 0);
        for (size_t q = 0; q < level.size(); q++) lev_ptr[(size_t) level[q]]++;  // counts at index level (1-based)
        // exclusive prefix: lev_ptr[l] = first position of level l+1
        uint16_t run = 0;
        for (int l = 1; l <= max_level; l++) {
            const uint16_t cnt = lev_ptr[(size_t) l];
            lev_ptr[(size_t) l - 1] = run;
            run = (uint16_t) (run + cnt);
        }
        lev_ptr[(size_t) max_level] = run;
        lev_bits.resize(level.size());
        std::vector<uint16_t> fill(lev_ptr.begin(), lev_ptr.end());
        for (size_t q = 0; q < level.size(); q++) lev_bits[fill[(size_t) level[q] - 1]++] = (uint16_t) h->serial_order[q];
        pl.n_levels = max_level;
        pl.mean_level = max_level ? (int) (level.size() / (size_t) max_level) : 0;
        pl.off_col_self = off;
        off += (uint32_t) (DVm * N);
        off = align_up(off, 4);
        pl.off_lev_ptr = off;
        off += 2u * (uint32_t) lev_ptr.size();
        off = align_up(off, 4);
        pl.off_lev_bits = off;
        off += 2u * (uint32_t) lev_bits.size();
    }
    off = align_up(off, 8);
    pl.off_prior = off;
    if (!h->uniform_prior) off += 8u * (uint32_t) g.n;
    off = align_up(off, 16);
    pl.blob.assign(off, 0);
    uint8_t *row_deg = pl.blob.data() + pl.off_row_deg;
    uint8_t *col_deg = pl.blob.data() + pl.off_col_deg;
    uint16_t *col_row = reinterpret_cast<uint16_t *>(pl.blob.data() + pl.off_col_row);
    uint16_t *row_pos = reinterpret_cast<uint16_t *>(pl.blob.data() + pl.off_row_pos);
    uint16_t *col_pos = reinterpret_cast<uint16_t *>(pl.blob.data() + pl.off_col_pos);
    // Message placement: conflict-free in both passes by edge colouring (place_messages above), 16 lanes per
    // shared-memory phase for 8-byte accesses.
    std::vector<uint32_t> slot_of_edge;
    const int longest = place_messages(g, 16, slot_of_edge) / 16;
    pl.msg_doubles = 16 * longest;
    if (pl.msg_doubles > 65535) {
        pl.why = "message positions do not fit 16-bit indices";
        return;
    }
    for (int i = 0; i < g.m; i++) {
        const uint32_t b = g.row_ptr[(size_t) i], e = g.row_ptr[(size_t) i + 1];
        row_deg[i] = (uint8_t) (e - b);
        for (uint32_t q = b; q < e; q++) {
            const uint32_t k = q - b;
            row_pos[2 * ((size_t) (k / 2) * M + i) + (k & 1)] = (uint16_t) slot_of_edge[q];
        }
    }
    if (pl.serial) {
        std::memcpy(pl.blob.data() + pl.off_lev_ptr, lev_ptr.data(), 2 * lev_ptr.size());
        std::memcpy(pl.blob.data() + pl.off_lev_bits, lev_bits.data(), 2 * lev_bits.size());
        uint8_t *col_self = pl.blob.data() + pl.off_col_self;
        for (int j = 0; j < g.n; j++)
            for (uint32_t q = g.col_ptr[(size_t) j]; q < g.col_ptr[(size_t) j + 1]; q++)
                col_self[(size_t) (q - g.col_ptr[(size_t) j]) * N + j] =
                    (uint8_t) (g.csc2csr[q] - g.row_ptr[g.row_idx[q]]);
    }
    for (int j = 0; j < g.n; j++) {
        const uint32_t b = g.col_ptr[(size_t) j], e = g.col_ptr[(size_t) j + 1];
        col_deg[j] = (uint8_t) (e - b);
        for (uint32_t q = b; q < e; q++) {
            col_pos[2 * ((size_t) ((q - b) / 2) * N + j) + ((q - b) & 1)] = (uint16_t) slot_of_edge[g.csc2csr[q]];
            col_row[2 * ((size_t) ((q - b) / 2) * N + j) + ((q - b) & 1)] = (uint16_t) g.row_idx[q];
        }
    }
    // verify: largest number of lanes of one half-warp access that share a bank pair (1 = conflict-free)
    pl.max_bank_multiplicity = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int items = pass == 0 ? g.m : g.n, stride = pass == 0 ? M : N, slots = pass == 0 ? DCm : DVm;
        const uint16_t *tab = pass == 0 ? row_pos : col_pos;
        const uint8_t *deg = pass == 0 ? row_deg : col_deg;
        for (int base = 0; base < items; base += 16)
            for (int k = 0; k < slots; k++) {
                int cnt[16] = {0};
                for (int x = base; x < std::min(items, base + 16); x++)
                    if (k < deg[x])
                        pl.max_bank_multiplicity = std::max(
                            pl.max_bank_multiplicity, ++cnt[tab[2 * ((size_t) (k / 2) * stride + x) + (k & 1)] & 15]);
            }
    }
    if (!h->uniform_prior) std::memcpy(pl.blob.data() + pl.off_prior, h->prior.data(), 8 * (size_t) g.n);
    uint32_t go = 0;
    pl.goff_msg = go;
    go += 8u * (uint32_t) pl.msg_doubles;
    pl.goff_dec = go;
    go += pl.serial ? (uint32_t) N : (uint32_t) N / 8;  // parallel: one bit per column; serial: one byte
    pl.goff_syn = go;
    go += 2u * 4u * (uint32_t) ((g.m + 31) / 32);  // packed syndrome + candidate accumulator
    go = align_up(go, 8);
    pl.goff_ctl = go;
    go += 8;
    pl.group_bytes = align_up(go, 16);
    if ((size_t) off + pl.group_bytes > (size_t) h->max_smem_optin) {
        pl.why = "one syndrome's messages do not fit in shared memory";
        return;
    }
    pl.ok = true;
}

// Tables of the paired on-chip family (bp_pair.cuh): parallel schedule only, double2 message slots placed by an
// 8-colouring (one colour per bank quad, 8 lanes per 16-byte shared-memory phase).
void build_pair_plan(bpb_decoder *h) {
    const bpb::HostGraph &g = h->g;
    bpb::PairPlan &pl = h->pair_plan;
    pl = bpb::PairPlan();
    const int DCm = g.max_row_degree, DVm = g.max_col_degree;
    const int M = (int) align_up((uint32_t) g.m, 32), N = (int) align_up((uint32_t) g.n, 32);
    if (DCm > 32 || DVm > 16 || (DCm > 8 && DVm > 4)) {
        pl.why = "degrees beyond the paired kernels' buckets";
        return;
    }
    if (DCm < 1 || g.n > 65535 || g.m > 65535) {
        pl.why = "row / column indices do not fit 16 bits";
        return;
    }
    std::vector<uint32_t> slot_of_edge;
    pl.msg_slots = place_messages(g, 8, slot_of_edge);
    if (pl.msg_slots > 65535) {
        pl.why = "message positions do not fit 16-bit indices";
        return;
    }
    pl.M = M;
    pl.N = N;
    const int DCp = (DCm + 1) / 2, DVp = (DVm + 1) / 2;
    uint32_t off = 0;
    pl.off_row_deg = off;
    off += (uint32_t) M;
    pl.off_col_deg = off;
    off += (uint32_t) N;
    pl.off_col_row = off;
    off += 4u * (uint32_t) (DVp * N);
    pl.off_row_pos = off;
    off += 4u * (uint32_t) (DCp * M);
    pl.off_col_pos = off;
    off += 4u * (uint32_t) (DVp * N);
    off = align_up(off, 8);
    pl.off_prior = off;
    if (!h->uniform_prior) off += 8u * (uint32_t) g.n;
    off = align_up(off, 16);
    pl.blob.assign(off, 0);
    uint8_t *row_deg = pl.blob.data() + pl.off_row_deg;
    uint8_t *col_deg = pl.blob.data() + pl.off_col_deg;
    uint16_t *col_row = reinterpret_cast<uint16_t *>(pl.blob.data() + pl.off_col_row);
    uint16_t *row_pos = reinterpret_cast<uint16_t *>(pl.blob.data() + pl.off_row_pos);
    uint16_t *col_pos = reinterpret_cast<uint16_t *>(pl.blob.data() + pl.off_col_pos);
    for (int i = 0; i < g.m; i++) {
        const uint32_t b = g.row_ptr[(size_t) i], e = g.row_ptr[(size_t) i + 1];
        row_deg[i] = (uint8_t) (e - b);
        for (uint32_t q = b; q < e; q++) {
            const uint32_t k = q - b;
            row_pos[2 * ((size_t) (k / 2) * M + i) + (k & 1)] = (uint16_t) slot_of_edge[q];
        }
    }
    for (int j = 0; j < g.n; j++) {
        const uint32_t b = g.col_ptr[(size_t) j], e = g.col_ptr[(size_t) j + 1];
        col_deg[j] = (uint8_t) (e - b);
        for (uint32_t q = b; q < e; q++) {
            col_pos[2 * ((size_t) ((q - b) / 2) * N + j) + ((q - b) & 1)] = (uint16_t) slot_of_edge[g.csc2csr[q]];
            col_row[2 * ((size_t) ((q - b) / 2) * N + j) + ((q - b) & 1)] = (uint16_t) g.row_idx[q];
        }
    }
    // verify: largest number of lanes of one quarter-warp access that share a bank quad (1 = conflict-free)
    for (int pass = 0; pass < 2; pass++) {
        const int items = pass == 0 ? g.m : g.n, stride = pass == 0 ? M : N, slots = pass == 0 ? DCm : DVm;
        const uint16_t *tab = pass == 0 ? row_pos : col_pos;
        const uint8_t *deg = pass == 0 ? row_deg : col_deg;
        for (int base = 0; base < items; base += 8)
            for (int k = 0; k < slots; k++) {
                int cnt[8] = {0};
                for (int x = base; x < std::min(items, base + 8); x++)
                    if (k < deg[x])
                        pl.max_bank_multiplicity = std::max(
                            pl.max_bank_multiplicity, ++cnt[tab[2 * ((size_t) (k / 2) * stride + x) + (k & 1)] & 7]);
            }
    }
    if (!h->uniform_prior) std::memcpy(pl.blob.data() + pl.off_prior, h->prior.data(), 8 * (size_t) g.n);
    const uint32_t MW = (uint32_t) ((g.m + 31) / 32);
    uint32_t go = 0;
    pl.goff_msg = go;
    go += 16u * (uint32_t) pl.msg_slots;
    pl.goff_syn = go;
    go += 2u * 4u * MW;  // two packed syndromes, interleaved word by word
    pl.goff_acc = go;
    go += 2u * 4u * MW;  // syndrome ^ candidate syndrome of the two halves, interleaved
    go = align_up(go, 8);
    pl.goff_ctl = go;
    go += 16;
    pl.group_bytes = align_up(go, 16);
    if ((size_t) off + pl.group_bytes > (size_t) h->max_smem_optin) {
        pl.why = "two syndromes' messages do not fit in shared memory";
        return;
    }
    pl.ok = true;
}

}  // namespace bpb
