// Host-side description of the proving system: tables, cross-table lookups, config, proof structures,
// transcript and wire format.  Mirrors (B200 side of) circuits/src/stark/{config,stark,cross_table_lookup,
// permutation,proof,serialization}.rs and plonky2/plonky2/src/iop/challenger.rs.
#pragma once
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gl.cuh"
#include "blake3.cuh"
#include "poseidon.cuh"

namespace ola {
namespace stark {

typedef uint64_t F;
typedef gl::ext2 E;

struct Hash {
    F e[4];
};
typedef std::vector<Hash> Cap;

// StarkConfig::standard_fast_config (circuits/src/stark/config.rs:18-30) -- the only configuration the
// reference ever instantiates; compile-time here.
struct Config {
    static constexpr uint32_t num_challenges = 2;
    static constexpr uint32_t rate_bits = 3;
    static constexpr uint32_t cap_height = 4;
    static constexpr uint32_t pow_bits = 16;
    static constexpr uint32_t arity_bits = 4;
    static constexpr uint32_t final_poly_bits = 5;
    static constexpr uint32_t num_queries = 28;
    bool check_quotient_degree = true;  // false: "pipeline parity" mode for non-satisfying synthetic traces
};

// FriReductionStrategy::ConstantArityBits(4, 5) (fri/reduction_strategies.rs:40-53)
inline std::vector<uint32_t> fri_arities(uint32_t degree_bits) {
    std::vector<uint32_t> r;
    uint32_t d = degree_bits;
    while (d > Config::final_poly_bits && d + Config::rate_bits - Config::arity_bits >= Config::cap_height) {
        r.push_back(Config::arity_bits);
        d -= Config::arity_bits;
    }
    return r;
}

// ---- Fiat-Shamir transcript (iop/challenger.rs:18-162): host side, exact and sequential ----
// Generic over H: H::Permutation is the Poseidon permutation or Blake3Permutation (hash/blake3.rs:165-199), and
// observe_hash absorbs hash.to_vec() -- 4 elements of a HashOut, 5 seven-byte elements of a BytesHash<32>
// (challenger.rs:79-81, hash/hash_types.rs:142-152).
// ---- an EXTERNAL transcript: the host application keeps its own Challenger (ola_prove_session_*) ----
// The prover runs on a worker thread and, instead of hashing, hands every transcript operation to the API caller as an
// event -- OBSERVE (elements to absorb, in order), CHALLENGE (n elements to squeeze), COMPACT (Challenger::compact) -- and
// blocks until the caller has consumed it (and, for CHALLENGE, supplied the values).  Stage labels say which seam of
// prove_with_traces / prove_single_table / fri_proof the event belongs to.
enum TranscriptStage {
    STAGE_TRACE_CAPS = 1,      // prover.rs:147-150: observe every table's trace cap
    STAGE_CTL_CHALLENGES = 2,  // cross_table_lookup_data's beta / gamma pairs (prover.rs:152-158)
    STAGE_TABLE_BEGIN = 3,     // prove_single_table: challenger.compact() + permutation challenges (prover.rs:344-372)
    STAGE_ZS_CAP = 4,          // permutation / CTL Z commitment (prover.rs:411-413)
    STAGE_ALPHAS = 5,          // get_n_challenges(num_challenges) (prover.rs:415)
    STAGE_QUOTIENT_CAP = 6,    // prover.rs:489-491
    STAGE_ZETA = 7,            // get_extension_challenge (prover.rs:493)
    STAGE_OPENINGS = 8,        // observe_openings (prover.rs:530)
    STAGE_FRI_ALPHA = 9,       // prove_openings: alpha (fri/oracle.rs:176)
    STAGE_FRI_LAYER_CAP = 10,  // fri_committed_trees: layer cap (fri/prover.rs:94)
    STAGE_FRI_BETA = 11,       // fri/prover.rs:96
    STAGE_FRI_FINAL_POLY = 12, // fri/prover.rs:118
    STAGE_FRI_POW = 13,        // fri_proof_of_work: challenger.get_hash() (fri/prover.rs:131)
    STAGE_FRI_QUERY_INDICES = 14  // fri_prover_query_rounds (fri/prover.rs:157-160)
};
struct TranscriptHost {
    enum Kind { NONE = 0, OBSERVE = 1, CHALLENGE = 2, COMPACT = 3, DONE = 4, FAILED = 5 };
    std::mutex m;
    std::condition_variable cv;
    Kind pending = NONE;     // event published by the prover, not yet consumed by the caller
    int stage = 0, table = -1;
    std::vector<F> elems;    // OBSERVE: the elements; CHALLENGE: filled by the caller
    size_t want = 0;         // CHALLENGE: how many
    bool supplied = false, consumed = false, abort = false;
    bool delivered = false;  // the caller has been handed the pending event (ola_prove_session_next)
    struct Aborted {};
    // prover side: publish an event and wait until the caller has dealt with it
    void publish(Kind k, int stg, int tbl, std::vector<F>&& data, size_t n_want) {
        std::unique_lock<std::mutex> lk(m);
        pending = k;
        stage = stg;
        table = tbl;
        elems = std::move(data);
        want = n_want;
        supplied = consumed = delivered = false;
        cv.notify_all();
        cv.wait(lk, [&] { return abort || (k == CHALLENGE ? supplied : consumed); });
        if (abort) throw Aborted();
        pending = NONE;
    }
};

struct Challenger {
    F state[12];
    std::vector<F> in, out;
    int hasher;  // OLA_HASH_POSEIDON (0) / OLA_HASH_BLAKE3 (1)
    TranscriptHost* host = nullptr;  // non-null: every operation is forwarded to the caller's Challenger instead
    int stage = 0, table = -1;       // labels of the forwarded events
    explicit Challenger(int hasher_id = 0, TranscriptHost* h = nullptr) : hasher(hasher_id), host(h) { memset(state, 0, sizeof(state)); }
    void at(int stg) {
        if (host && stg != stage) flush_to_host();  // elements observed under the previous label are delivered with it
        stage = stg;
    }
    void flush_to_host() {
        if (!in.empty()) {
            std::vector<F> data;
            data.swap(in);
            host->publish(TranscriptHost::OBSERVE, stage, table, std::move(data), 0);
        }
    }
    std::vector<F> ask_host(size_t n) {
        flush_to_host();
        host->publish(TranscriptHost::CHALLENGE, stage, table, std::vector<F>(), n);
        std::vector<F> r;
        {
            std::lock_guard<std::mutex> lk(host->m);
            r = host->elems;
        }
        if (r.size() != n) throw std::runtime_error("transcript host supplied a wrong number of challenges");
        for (auto& x : r) x = gl::canon(x);
        return r;
    }
    void duplexing() {
        for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
        in.clear();
        if (hasher == 1)
            blake3::permute_host(state);
        else
            poseidon::permute_host(state);
        out.assign(state, state + 8);
    }
    void observe(F x) {
        if (host) {
            in.push_back(gl::canon(x));  // buffered: forwarded as one OBSERVE event before the next challenge
            return;
        }
        out.clear();
        in.push_back(gl::canon(x));
        if (in.size() == 8) duplexing();
    }
    void observe_ext(E x) {
        observe(x.c0);
        observe(x.c1);
    }
    void observe_cap(const Cap& c) {
        for (auto& h : c) {
            if (hasher == 1) {
                F f[5];
                blake3::hash_to_fields(h.e, f);
                for (int i = 0; i < 5; i++) observe(f[i]);
            } else {
                for (int i = 0; i < 4; i++) observe(h.e[i]);
            }
        }
    }
    F get_challenge() {
        if (host) return ask_host(1)[0];
        if (!in.empty() || out.empty()) duplexing();
        F r = out.back();  // pops from the END (challenger.rs:97-99)
        out.pop_back();
        return r;
    }
    std::vector<F> get_challenges(size_t n) {  // get_n_challenges (challenger.rs:101-103)
        if (host) return ask_host(n);
        std::vector<F> r;
        for (size_t i = 0; i < n; i++) r.push_back(get_challenge());
        return r;
    }
    E get_ext() {
        if (host) {
            std::vector<F> r = ask_host(2);
            return gl::make2(r[0], r[1]);
        }
        F a = get_challenge();
        F b = get_challenge();
        return gl::make2(a, b);
    }
    Hash get_hash() {
        if (host) {
            std::vector<F> r = ask_host(4);
            Hash hh;
            for (int i = 0; i < 4; i++) hh.e[i] = r[i];
            return hh;
        }
        Hash h;
        for (int i = 0; i < 4; i++) h.e[i] = get_challenge();
        return h;
    }
    void compact() {
        if (host) {
            flush_to_host();
            host->publish(TranscriptHost::COMPACT, stage, table, std::vector<F>(), 0);
            return;
        }
        if (!in.empty()) duplexing();
        out.clear();
    }
};

// ---- cross-table lookups (cross_table_lookup.rs:29-170) ----
struct Column {
    std::vector<std::pair<int, F>> lc;
    F constant = 0;
    static Column single(int c) {
        Column r;
        r.lc.push_back({c, 1});
        return r;
    }
    static Column linear(std::vector<std::pair<int, F>> v, F k = 0) {
        Column r;
        r.lc = std::move(v);
        r.constant = k;
        return r;
    }
};
inline std::vector<Column> singles(std::initializer_list<int> cs) {
    std::vector<Column> r;
    for (int c : cs) r.push_back(Column::single(c));
    return r;
}
struct TableWithColumns {
    int table = 0;
    std::vector<Column> columns;
    bool has_filter = false;
    Column filter;
};
inline TableWithColumns twc(int table, std::vector<Column> cols, Column filter) {
    TableWithColumns t;
    t.table = table;
    t.columns = std::move(cols);
    t.has_filter = true;
    t.filter = std::move(filter);
    return t;
}
struct CrossTableLookup {
    std::vector<TableWithColumns> looking;
    TableWithColumns looked;
    bool has_looked = true;
    bool missing_sides = false;  // a side's table has no constraint kernel in this build yet
    bool complete = true;  // every side of the registered CTL is inside the system
};
struct Challenge {
    F beta, gamma;
};
struct PermutationPair {
    std::vector<std::pair<int, int>> column_pairs;
};

// reference Table enum (ola_stark.rs:104-119)
enum TableId { T_CPU = 0, T_MEMORY, T_BITWISE, T_CMP, T_RANGECHECK, T_POSEIDON, T_POSEIDON_CHUNK, T_STORAGE, T_TAPE, T_SCCALL, T_PROGRAM, T_PROG_CHUNK, T_NUM };

struct TableInfo {
    int id = -1;
    const char* name = "";
    int columns = 0;
    int constraint_degree = 0;
    std::vector<PermutationPair> permutation_pairs;
    F compress_challenge = 0;  // Bitwise / Program: beta from trace generation
    int quotient_degree_factor() const { return constraint_degree - 1 > 1 ? constraint_degree - 1 : 1; }
    int permutation_batch_size() const { return quotient_degree_factor(); }
    int num_permutation_batches() const {
        int inst = (int)permutation_pairs.size() * (int)Config::num_challenges;
        int bs = permutation_batch_size();
        return (inst + bs - 1) / bs;
    }
};
struct System {
    std::vector<TableInfo> tables;
    std::vector<CrossTableLookup> ctls;  // table fields = positions inside `tables`
    std::vector<F> compress_challenges;
};
// registry (air/registry.cuh): metadata for the tables whose constraint kernels are compiled in
bool table_available(int id);
TableInfo table_info(int id);
std::vector<CrossTableLookup> all_cross_table_lookups();
System make_system(const std::vector<int>& ids);

// ---- proof structures (proof.rs:108-119, :181-196; fri/proof.rs:105-114) ----
struct FriQueryStep {
    std::vector<E> evals;
    std::vector<Hash> siblings;
};
struct FriQueryRound {
    std::vector<std::pair<std::vector<F>, std::vector<Hash>>> initial;
    std::vector<FriQueryStep> steps;
};
struct FriProof {
    std::vector<Cap> commit_caps;
    std::vector<FriQueryRound> rounds;
    std::vector<E> final_poly;
    F pow_witness = 0;
};
struct OpeningSet {
    std::vector<E> local_values, next_values, zs, zs_next;
    std::vector<F> ctl_zs_last;
    std::vector<E> quotient;
};
struct StarkProof {
    Cap trace_cap, zs_cap, quotient_cap;
    OpeningSet openings;
    FriProof fri;
};

// ---- wire format: Buffer::write_all_proof (serialization.rs:349-393): LE canonical u64, u32 length prefixes,
// u8 Merkle-path length; PublicValues are not written ----
struct Writer {
    std::vector<uint8_t> buf;
    bool raw_hashes = false;  // BytesHash<32>: write_hash = the 32 digest bytes (serialization.rs:115-117), never reduced
    void u8(uint8_t x) { buf.push_back(x); }
    void u32(uint32_t x) {
        for (int i = 0; i < 4; i++) buf.push_back((uint8_t)(x >> (8 * i)));
    }
    void field(F x) {
        x = gl::canon(x);
        for (int i = 0; i < 8; i++) buf.push_back((uint8_t)(x >> (8 * i)));
    }
    void ext(E x) {
        field(x.c0);
        field(x.c1);
    }
    void field_vec(const std::vector<F>& v) {
        u32((uint32_t)v.size());
        for (F x : v) field(x);
    }
    void ext_vec(const std::vector<E>& v) {
        u32((uint32_t)v.size());
        for (E x : v) ext(x);
    }
    void hash(const Hash& h) {
        for (int i = 0; i < 4; i++) {
            if (raw_hashes)
                for (int k = 0; k < 8; k++) buf.push_back((uint8_t)(h.e[i] >> (8 * k)));
            else
                field(h.e[i]);
        }
    }
    void cap(const Cap& c) {
        u32((uint32_t)c.size());
        for (auto& h : c) hash(h);
    }
    void merkle_proof(const std::vector<Hash>& s) {
        u8((uint8_t)s.size());
        for (auto& h : s) hash(h);
    }
    void proof(const StarkProof& p) {
        cap(p.trace_cap);
        cap(p.zs_cap);
        cap(p.quotient_cap);
        ext_vec(p.openings.local_values);
        ext_vec(p.openings.next_values);
        ext_vec(p.openings.zs);
        ext_vec(p.openings.zs_next);
        field_vec(p.openings.ctl_zs_last);
        ext_vec(p.openings.quotient);
        u32((uint32_t)p.fri.commit_caps.size());
        for (auto& c : p.fri.commit_caps) cap(c);
        u32((uint32_t)p.fri.rounds.size());
        for (auto& r : p.fri.rounds) {
            u32((uint32_t)r.initial.size());
            for (auto& ip : r.initial) {
                field_vec(ip.first);
                merkle_proof(ip.second);
            }
            u32((uint32_t)r.steps.size());
            for (auto& s : r.steps) {
                ext_vec(s.evals);
                merkle_proof(s.siblings);
            }
        }
        ext_vec(p.fri.final_poly);
        field(p.fri.pow_witness);
    }
};

}  // namespace stark
}  // namespace ola
