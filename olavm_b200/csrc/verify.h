// Proof verification (host code; no GPU needed): the product-side counterpart of
//   circuits/src/stark/verifier.rs      verify_proof :32-212, verify_stark_proof_with_challenges :214-300,
//                                       validate_proof_shape :310-360, eval_l_0_and_l_last :380-396
//   circuits/src/stark/get_challenges.rs  AllProof::get_challenges :19-75, StarkProof::get_challenges :129-208
//   circuits/src/stark/cross_table_lookup.rs  CtlCheckVars::from_proofs :337-377, verify_cross_table_lookups :560-600
//   circuits/src/stark/serialization.rs   Buffer::read_all_proof :395-411
//   plonky2/plonky2/src/fri/verifier.rs   verify_fri_proof :59-115, fri_combine_initial :117-160,
//                                       fri_verifier_query_round :162-232, compute_evaluation :18-41
//   plonky2/plonky2/src/fri/challenges.rs fri_challenges :25-75;  hash/merkle_proofs.rs verify_merkle_proof_to_cap :36-64
// The constraint bodies are the same transcriptions the quotient kernels compile (csrc/air/*.h), instantiated over the
// quadratic extension at the opening point.  SURVEY.md 8(f) rank 2: "the only acceptance test the reference has".
#pragma once
#include <string>
#include <vector>

#include "air/registry.cuh"
#include "stark_types.h"

namespace ola {
namespace stark {
namespace verify {

// (host-only in practice; marked for both sides so that the shared __host__ __device__ constraint templates instantiate cleanly)
#if defined(__CUDACC__)
#define OLA_VHD __host__ __device__
#else
#define OLA_VHD
#endif

// ---- field value type for AIR evaluation at zeta ----
struct XE {
    E v;
    OLA_VHD XE() : v(gl::make2(0, 0)) {}
    OLA_VHD explicit XE(E x) : v(x) {}
    OLA_VHD explicit XE(F x) : v(gl::make2(gl::canon(x), 0)) {}
    OLA_VHD XE operator+(XE o) const { return XE(gl::add(v, o.v)); }
    OLA_VHD XE operator-(XE o) const { return XE(gl::sub(v, o.v)); }
    OLA_VHD XE operator*(XE o) const { return XE(gl::mul(v, o.v)); }
};
struct XRow {
    const XE* p;
    OLA_VHD XE operator[](int c) const { return p[c]; }
};
struct XConsumer {  // ConstraintConsumer over the extension (constraint_consumer.rs:10-80)
    XE alpha[2], acc[2], z_last, lagrange_first, lagrange_last;
    OLA_VHD void constraint(XE c) {
        for (int j = 0; j < 2; ++j) acc[j] = acc[j] * alpha[j] + c;
    }
    OLA_VHD void constraint_transition(XE c) { constraint(c * z_last); }
    OLA_VHD void constraint_first_row(XE c) { constraint(c * lagrange_first); }
    OLA_VHD void constraint_last_row(XE c) { constraint(c * lagrange_last); }
};
struct HostPoseidonParams {  // parameter tables for the Poseidon table's AIR (host copies)
    static OLA_VHD uint64_t round(int i) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return gl::canon(OLA_ALL_ROUND_CONSTANTS[i]);
#endif
    }
    static OLA_VHD uint64_t circ(int i) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return OLA_MDS_MATRIX_CIRC[i];
#endif
    }
    static OLA_VHD uint64_t diag(int i) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return OLA_MDS_MATRIX_DIAG[i];
#endif
    }
    static OLA_VHD uint64_t first(int i) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return gl::canon(OLA_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
#endif
    }
    static OLA_VHD uint64_t partial(int r) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return gl::canon(OLA_FAST_PARTIAL_ROUND_CONSTANTS[r]);
#endif
    }
    static OLA_VHD uint64_t init(int r, int c) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return gl::canon(OLA_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r][c]);
#endif
    }
    static OLA_VHD uint64_t what(int r, int i) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return gl::canon(OLA_FAST_PARTIAL_ROUND_W_HATS[r][i]);
#endif
    }
    static OLA_VHD uint64_t vs(int r, int i) {
#if defined(__CUDA_ARCH__)
        return 0;  // never called on the device
#else
        return gl::canon(OLA_FAST_PARTIAL_ROUND_VS[r][i]);
#endif
    }
};

}  // namespace verify
}  // namespace stark
namespace air {
template <>
OLA_VHD inline stark::verify::XE kc<stark::verify::XE>(uint64_t k) {
    return stark::verify::XE((stark::F)k);
}
template <>
OLA_VHD inline bool is_zero<stark::verify::XE>(const stark::verify::XE& x) {
    return gl::canon(x.v.c0) == 0 && gl::canon(x.v.c1) == 0;
}
}  // namespace air
namespace stark {
namespace verify {

// the table's eval_packed_generic at (local, next) over the extension
template <class YC>
inline void eval_table_t(const TableInfo& t, const XRow& lv, const XRow& nv, YC& yc) {
    const XE beta(t.compress_challenge);
    switch (t.id) {
        case T_CPU: air::cpu::eval<XE, XRow, YC>(lv, nv, yc); break;
        case T_MEMORY: air::mem::eval<XE, XRow, YC>(lv, nv, yc); break;
        case T_BITWISE: air::bitwise::eval<XE, XRow, YC>(lv, nv, yc, beta); break;
        case T_CMP: {  // cmp_stark.rs:21-45
            const XE one((F)1), op0 = lv[0], op1 = lv[1], gte = lv[2], abs_diff = lv[3], abs_diff_inv = lv[4];
            yc.constraint(gte * (one - gte));
            yc.constraint(gte * (op0 - op1 - abs_diff));
            yc.constraint((one - gte) * (op1 - op0 - abs_diff));
            yc.constraint((one - gte) * (one - abs_diff * abs_diff_inv));
            break;
        }
        case T_RANGECHECK: {  // rangecheck_stark.rs:27-67
            const XE val = lv[4], limb_lo = lv[5], limb_hi = lv[6];
            yc.constraint(val - (limb_lo + limb_hi * XE((F)(1 << 16))));
            air::eval_lookups_t<XE, XRow, YC>(lv, nv, yc, 7, 10);
            air::eval_lookups_t<XE, XRow, YC>(lv, nv, yc, 8, 11);
            break;
        }
        case T_POSEIDON: air::psdn::eval<XE, XRow, YC, HostPoseidonParams>(lv, nv, yc); break;
        case T_POSEIDON_CHUNK: air::psdn_chunk::eval<XE, XRow, YC>(lv, nv, yc); break;
        case T_STORAGE: air::storage::eval<XE, XRow, YC>(lv, nv, yc); break;
        case T_TAPE: air::tape::eval<XE, XRow, YC>(lv, nv, yc); break;
        case T_SCCALL: air::sccall::eval<XE, XRow, YC>(lv, nv, yc); break;
        case T_PROGRAM: air::program::eval<XE, XRow, YC>(lv, nv, yc, beta); break;
        case T_PROG_CHUNK: air::prog_chunk::eval<XE, XRow, YC>(lv, nv, yc); break;
        default: throw Error(OLA_ERR_INVALID_ARG, "unknown table id");
    }
}

inline void eval_table(const TableInfo& t, const XRow& lv, const XRow& nv, XConsumer& yc) { eval_table_t<XConsumer>(t, lv, nv, yc); }
// number of constraints the table's eval_packed_generic emits (independent of the row: no constraint is conditional)
struct CountingConsumer {
    int n = 0;
    void constraint(XE) { ++n; }
    void constraint_transition(XE) { ++n; }
    void constraint_first_row(XE) { ++n; }
    void constraint_last_row(XE) { ++n; }
};
// the raw argument and the kind (0 constraint, 1 transition, 2 first row, 3 last row) of every yield, in order
struct RecordingConsumer {
    std::vector<XE> vals;
    std::vector<int> kinds;
    void constraint(XE c) { vals.push_back(c); kinds.push_back(0); }
    void constraint_transition(XE c) { vals.push_back(c); kinds.push_back(1); }
    void constraint_first_row(XE c) { vals.push_back(c); kinds.push_back(2); }
    void constraint_last_row(XE c) { vals.push_back(c); kinds.push_back(3); }
};
inline int air_constraint_count(const TableInfo& t) {
    std::vector<XE> zeros((size_t)t.columns, XE((F)0));
    XRow row{zeros.data()};
    CountingConsumer cc;
    eval_table_t<CountingConsumer>(t, row, row, cc);
    return cc.n;
}

// C::Hasher the proof under verification was made with (set by verify_all for its duration; one per host thread)
inline int& current_hasher() {
    static thread_local int h = 0;
    return h;
}
struct HasherScope {
    int prev;
    explicit HasherScope(int h) : prev(current_hasher()) { current_hasher() = h; }
    ~HasherScope() { current_hasher() = prev; }
};

// ---- wire format reader (the inverse of Writer, stark_types.h) ----
struct Reader {
    const uint8_t* p;
    size_t len, pos = 0;
    bool ok = true;
    Reader(const uint8_t* d, size_t n) : p(d), len(n) {}
    bool need(size_t k) {
        if (!ok || len - pos < k) ok = false;
        return ok;
    }
    uint8_t u8() { return need(1) ? p[pos++] : 0; }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t x = 0;
        for (int i = 0; i < 4; ++i) x |= (uint32_t)p[pos++] << (8 * i);
        return x;
    }
    F field() {
        if (!need(8)) return 0;
        uint64_t x = 0;
        for (int i = 0; i < 8; ++i) x |= (uint64_t)p[pos++] << (8 * i);
        if (x >= gl::P) ok = false;  // read_field: canonical encodings only
        return x;
    }
    E ext() {
        F a = field(), b = field();
        return gl::make2(a, b);
    }
    size_t count(size_t elem_bytes) {  // a u32 length prefix that the remaining bytes can actually hold
        uint32_t k = u32();
        if (ok && (size_t)k * elem_bytes > len - pos) ok = false;
        return ok ? k : 0;
    }
    std::vector<F> field_vec() {
        size_t k = count(8);
        std::vector<F> v(k);
        for (auto& x : v) x = field();
        return v;
    }
    std::vector<E> ext_vec() {
        size_t k = count(16);
        std::vector<E> v(k);
        for (auto& x : v) x = ext();
        return v;
    }
    Hash hash() {
        Hash h;
        if (current_hasher() == 1) {  // BytesHash<32>::from_bytes: any 32 bytes
            for (int i = 0; i < 4; ++i) {
                h.e[i] = 0;
                if (need(8))
                    for (int k = 0; k < 8; ++k) h.e[i] |= (uint64_t)p[pos++] << (8 * k);
            }
            return h;
        }
        for (int i = 0; i < 4; ++i) h.e[i] = field();
        return h;
    }
    Cap cap() {
        size_t k = count(32);
        Cap c(k);
        for (auto& h : c) h = hash();
        return c;
    }
    std::vector<Hash> merkle_proof() {
        size_t k = u8();
        if (ok && k * 32 > len - pos) ok = false;
        std::vector<Hash> v(ok ? k : 0);
        for (auto& h : v) h = hash();
        return v;
    }
    StarkProof proof() {
        StarkProof q;
        q.trace_cap = cap();
        q.zs_cap = cap();
        q.quotient_cap = cap();
        q.openings.local_values = ext_vec();
        q.openings.next_values = ext_vec();
        q.openings.zs = ext_vec();
        q.openings.zs_next = ext_vec();
        q.openings.ctl_zs_last = field_vec();
        q.openings.quotient = ext_vec();
        size_t nc = count(4);
        for (size_t i = 0; i < nc && ok; ++i) q.fri.commit_caps.push_back(cap());
        size_t nr = count(4);
        for (size_t i = 0; i < nr && ok; ++i) {
            FriQueryRound r;
            size_t ni = count(5);
            for (size_t k = 0; k < ni && ok; ++k) {
                std::vector<F> row = field_vec();
                r.initial.push_back({row, merkle_proof()});
            }
            size_t ns = count(5);
            for (size_t k = 0; k < ns && ok; ++k) {
                FriQueryStep s;
                s.evals = ext_vec();
                s.siblings = merkle_proof();
                r.steps.push_back(std::move(s));
            }
            q.fri.rounds.push_back(std::move(r));
        }
        q.fri.final_poly = ext_vec();
        q.fri.pow_witness = field();
        return q;
    }
};

// ---- hashing on the host (hashing.rs:84-108 hash_n_to_hash_no_pad, :66-74 two_to_one) ----
inline Hash poseidon_hash_no_pad(const F* in, size_t n) {
    F st[12] = {0};
    for (size_t i = 0; i < n; i += 8) {
        for (size_t k = 0; k < 8 && i + k < n; ++k) st[k] = in[i + k];
        poseidon::permute_host(st);
    }
    Hash h;
    for (int i = 0; i < 4; ++i) h.e[i] = st[i];
    return h;
}
inline Hash hash_no_pad(const F* in, size_t n) {
    if (current_hasher() == 1) {  // Blake3_256::hash_no_pad (hash/blake3.rs:205-218)
        Hash h;
        if (n > blake3::MAX_U64S) throw Error(OLA_ERR_INVALID_ARG, "leaf too wide");
        std::vector<F> c(in, in + n);
        for (auto& x : c) x = gl::canon(x);
        blake3::hash_host(c.data(), n, h.e);
        return h;
    }
    return poseidon_hash_no_pad(in, n);
}
inline Hash two_to_one(const Hash& l, const Hash& r) {
    if (current_hasher() == 1) {  // Blake3_256::two_to_one (hash/blake3.rs:220-233)
        Hash h;
        blake3::two_to_one(l.e, r.e, h.e);
        return h;
    }
    F st[12] = {l.e[0], l.e[1], l.e[2], l.e[3], r.e[0], r.e[1], r.e[2], r.e[3], 0, 0, 0, 0};
    poseidon::permute_host(st);
    Hash h;
    for (int i = 0; i < 4; ++i) h.e[i] = st[i];
    return h;
}
inline bool merkle_verify(const F* leaf, size_t nleaf, size_t index, const Cap& cap, const std::vector<Hash>& sib) {
    Hash cur = hash_no_pad(leaf, nleaf);  // the fork hashes every leaf (merkle_proofs.rs:60, merkle_tree/mod.rs:198), no hash_or_noop
    for (auto& s : sib) {
        cur = (index & 1) ? two_to_one(s, cur) : two_to_one(cur, s);
        index >>= 1;
    }
    if (index >= cap.size()) return false;
    for (int i = 0; i < 4; ++i) {
        if (current_hasher() == 1 ? cur.e[i] != cap[index].e[i] : gl::canon(cur.e[i]) != gl::canon(cap[index].e[i])) return false;
    }
    return true;
}

inline E efrom(F x) { return gl::make2(gl::canon(x), 0); }
inline bool eeq(E a, E b) { return gl::canon(a.c0) == gl::canon(b.c0) && gl::canon(a.c1) == gl::canon(b.c1); }
inline size_t bitrev(size_t x, uint32_t bits) {
    size_t r = 0;
    for (uint32_t i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

struct OpeningBatch {
    E point;
    std::vector<std::pair<int, int>> polys;  // (oracle, column)
};
struct FriChallenges {
    E alpha;
    std::vector<E> betas;
    F pow_response = 0;
    std::vector<size_t> indices;
};

inline FriChallenges fri_challenges(Challenger& ch, const FriProof& p, uint32_t degree_bits) {
    FriChallenges fc;
    fc.alpha = ch.get_ext();
    for (auto& cap : p.commit_caps) {
        ch.observe_cap(cap);
        fc.betas.push_back(ch.get_ext());
    }
    for (auto& x : p.final_poly) ch.observe_ext(x);
    Hash h = ch.get_hash();
    F in[5] = {h.e[0], h.e[1], h.e[2], h.e[3], p.pow_witness};
    fc.pow_response = gl::canon(poseidon_hash_no_pad(in, 5).e[0]);  // C::InnerHasher = PoseidonHash in every config
    const size_t L = (size_t)1 << (degree_bits + Config::rate_bits);
    for (uint32_t i = 0; i < Config::num_queries; ++i) fc.indices.push_back((size_t)(ch.get_challenge() % L));
    return fc;
}

// value at x of the polynomial through (xs[i], ys[i]) (interpolation.rs barycentric form, evaluated directly)
inline E interpolate_eval(const std::vector<E>& xs, const std::vector<E>& ys, E x) {
    E r = gl::make2(0, 0);
    for (size_t i = 0; i < xs.size(); ++i) {
        E num = ys[i], den = gl::make2(1, 0);
        for (size_t j = 0; j < xs.size(); ++j)
            if (j != i) {
                num = gl::mul(num, gl::sub(x, xs[j]));
                den = gl::mul(den, gl::sub(xs[i], xs[j]));
            }
        r = gl::add(r, gl::mul(num, gl::inv(den)));
    }
    return r;
}

inline std::string verify_fri(const std::vector<OpeningBatch>& batches, const std::vector<size_t>& oracle_cols, const std::vector<std::vector<E>>& openings,
                              const FriChallenges& fc, const std::vector<const Cap*>& initial_caps, const FriProof& p, uint32_t degree_bits) {
    const uint32_t lde_bits = degree_bits + Config::rate_bits;
    const std::vector<uint32_t> arities = fri_arities(degree_bits);
    uint32_t total_ar = 0;
    for (auto a : arities) total_ar += a;
    if ((fc.pow_response >> (64 - Config::pow_bits)) != 0) return "Invalid proof of work witness.";
    if (p.rounds.size() != Config::num_queries) return "Number of query rounds does not match config.";
    if (p.commit_caps.size() != arities.size()) return "Number of FRI commit caps does not match the reduction strategy.";
    if (p.final_poly.size() != ((size_t)1 << (degree_bits - total_ar))) return "Final polynomial has the wrong length.";
    std::vector<E> reduced;  // PrecomputedReducedOpenings: reduce_with_powers(openings, alpha)
    for (auto& b : openings) {
        E s = gl::make2(0, 0);
        for (size_t i = b.size(); i-- > 0;) s = gl::add(gl::mul(s, fc.alpha), b[i]);
        reduced.push_back(s);
    }
    for (size_t qi = 0; qi < p.rounds.size(); ++qi) {
        size_t x = fc.indices[qi];
        const FriQueryRound& qr = p.rounds[qi];
        if (qr.initial.size() != initial_caps.size()) return "Wrong number of initial-tree openings.";
        for (size_t o = 0; o < qr.initial.size(); ++o) {
            if (qr.initial[o].first.size() != oracle_cols[o]) return "Initial-tree leaf has the wrong width.";
            if (qr.initial[o].second.size() + Config::cap_height != lde_bits) return "Initial-tree Merkle path has the wrong length.";
            if (!merkle_verify(qr.initial[o].first.data(), qr.initial[o].first.size(), x, *initial_caps[o], qr.initial[o].second)) return "Invalid Merkle proof.";
        }
        F sx = gl::mul(gl::GEN, gl::pow(gl::root_of_unity((int)lde_bits), (uint64_t)bitrev(x, lde_bits)));
        // fri_combine_initial: sum over batches of (reduced(leaf) - reduced(openings)) / (x - z), alpha-shifted
        E sum = gl::make2(0, 0);
        const E subgroup_x = efrom(sx);
        for (size_t bi = 0; bi < batches.size(); ++bi) {
            const OpeningBatch& b = batches[bi];
            E red = gl::make2(0, 0);
            for (size_t i = b.polys.size(); i-- > 0;) red = gl::add(gl::mul(red, fc.alpha), efrom(qr.initial[b.polys[i].first].first[b.polys[i].second]));
            const E num = gl::sub(red, reduced[bi]), den = gl::sub(subgroup_x, b.point);
            sum = gl::mul(sum, gl::pow(fc.alpha, (uint64_t)b.polys.size()));
            sum = gl::add(sum, gl::mul(num, gl::inv(den)));
        }
        E old_eval = gl::mul(sum, subgroup_x);
        if (qr.steps.size() != arities.size()) return "Wrong number of FRI query steps.";
        uint32_t bits = lde_bits;
        for (size_t i = 0; i < arities.size(); ++i) {
            const uint32_t ab = arities[i];
            const size_t arity = (size_t)1 << ab;
            const std::vector<E>& ev = qr.steps[i].evals;
            if (ev.size() != arity) return "FRI query step has the wrong number of evaluations.";
            const size_t coset_index = x >> ab, within = x & (arity - 1);
            if (!eeq(ev[within], old_eval)) return "FRI consistency check failed.";
            // compute_evaluation: interpolate the coset's values (bit-reversed order) and evaluate at beta
            const F g = gl::root_of_unity((int)ab);
            std::vector<E> evs(arity), xs;
            for (size_t k = 0; k < arity; ++k) evs[bitrev(k, ab)] = ev[k];
            const F start = gl::mul(sx, gl::pow(g, (uint64_t)(arity - bitrev(within, ab))));
            F y = 1;
            for (size_t k = 0; k < arity; ++k) {
                xs.push_back(efrom(gl::mul(start, y)));
                y = gl::mul(y, g);
            }
            old_eval = interpolate_eval(xs, evs, fc.betas[i]);
            std::vector<F> flat;
            for (auto& e : ev) {
                flat.push_back(e.c0);
                flat.push_back(e.c1);
            }
            bits -= ab;
            if (qr.steps[i].siblings.size() + Config::cap_height != bits) return "FRI step Merkle path has the wrong length.";
            if (!merkle_verify(flat.data(), flat.size(), coset_index, p.commit_caps[i], qr.steps[i].siblings)) return "Invalid Merkle proof.";
            for (uint32_t k = 0; k < ab; ++k) sx = gl::mul(sx, sx);
            x = coset_index;
        }
        E fin = gl::make2(0, 0);  // final_poly.eval(subgroup_x)
        for (size_t k = p.final_poly.size(); k-- > 0;) fin = gl::add(gl::mul(fin, efrom(sx)), p.final_poly[k]);
        if (!eeq(fin, old_eval)) return "Final polynomial evaluation is invalid.";
    }
    return "";
}

struct CtlVars {  // CtlCheckVars (cross_table_lookup.rs:325-335)
    E local_z, next_z;
    Challenge ch;
    const TableWithColumns* twc;
};
inline XE eval_column(const Column& c, const XRow& row) {
    XE s((F)c.constant);
    for (auto& t : c.lc) s = s + row[t.first] * XE(t.second);
    return s;
}

// verify_proof over the given system; returns "" when the proof is accepted, else the reason
inline std::string verify_all(const uint8_t* bytes, size_t len, const std::vector<int>& table_ids, int hasher = 0) {
    HasherScope scope(hasher);
    Reader rd(bytes, len);
    const size_t T = rd.count(1);
    if (!rd.ok || T != table_ids.size()) return "wrong number of proofs";
    std::vector<StarkProof> proofs;
    for (size_t i = 0; i < T && rd.ok; ++i) proofs.push_back(rd.proof());
    std::vector<F> cc = rd.field_vec();
    if (!rd.ok || rd.pos != len || cc.size() != T) return "malformed proof bytes";
    System sys = make_system(table_ids);
    for (size_t i = 0; i < T; ++i)  // verifier.rs:78-86: the compress challenges come from the proof
        if (sys.tables[i].id == T_BITWISE || sys.tables[i].id == T_PROGRAM) sys.tables[i].compress_challenge = cc[i];
    for (auto& p : proofs) {
        const size_t ncap = (size_t)1 << Config::cap_height;
        if (p.trace_cap.size() != ncap || p.zs_cap.size() != ncap || p.quotient_cap.size() != ncap) return "cap shape";
        for (auto& c : p.fri.commit_caps)
            if (c.size() != ncap) return "cap shape";
    }
    Challenger ch(current_hasher());
    for (auto& p : proofs) ch.observe_cap(p.trace_cap);
    std::vector<Challenge> ctl_ch;
    for (uint32_t k = 0; k < Config::num_challenges; ++k) {
        F b = ch.get_challenge();
        F g = ch.get_challenge();
        ctl_ch.push_back({b, g});
    }
    std::vector<size_t> nperm(T), cursor(T, 0);
    for (size_t i = 0; i < T; ++i) nperm[i] = (size_t)sys.tables[i].num_permutation_batches();
    std::vector<std::vector<CtlVars>> ctl_vars(T);
    for (auto& ctl : sys.ctls)
        for (auto& c : ctl_ch) {
            auto push = [&](const TableWithColumns& tw) {
                const StarkProof& p = proofs[tw.table];
                const size_t k = nperm[tw.table] + cursor[tw.table]++;
                if (k >= p.openings.zs.size() || k >= p.openings.zs_next.size()) return false;
                ctl_vars[tw.table].push_back({p.openings.zs[k], p.openings.zs_next[k], c, &tw});
                return true;
            };
            for (auto& lt : ctl.looking)
                if (!push(lt)) return "cross-table lookup openings missing";
            if (ctl.has_looked && !push(ctl.looked)) return "cross-table lookup openings missing";
        }
    for (size_t i = 0; i < T; ++i) {
        const TableInfo& t = sys.tables[i];
        const StarkProof& p = proofs[i];
        const std::string name = t.name;
        ch.compact();
        // recover_degree_bits (proof.rs:121-131)
        if (p.fri.rounds.empty() || p.fri.rounds[0].initial.empty()) return name + ": empty FRI proof";
        const uint32_t lde_bits = (uint32_t)p.fri.rounds[0].initial[0].second.size() + Config::cap_height;
        if (lde_bits < Config::rate_bits + 1 || lde_bits > 32) return name + ": degree out of range";
        const uint32_t degree_bits = lde_bits - Config::rate_bits;
        std::vector<std::vector<Challenge>> perm_sets;
        if (!t.permutation_pairs.empty())
            for (int s = 0; s < t.permutation_batch_size(); ++s) {
                std::vector<Challenge> set;
                for (uint32_t k = 0; k < Config::num_challenges; ++k) {
                    F b = ch.get_challenge();
                    F g = ch.get_challenge();
                    set.push_back({b, g});
                }
                perm_sets.push_back(set);
            }
        ch.observe_cap(p.zs_cap);
        const F alpha0 = ch.get_challenge(), alpha1 = ch.get_challenge();
        ch.observe_cap(p.quotient_cap);
        const E zeta = ch.get_ext();
        const OpeningSet& os = p.openings;
        for (auto& v : os.local_values) ch.observe_ext(v);
        for (auto& v : os.zs) ch.observe_ext(v);
        for (auto& v : os.quotient) ch.observe_ext(v);
        for (auto& v : os.next_values) ch.observe_ext(v);
        for (auto& v : os.zs_next) ch.observe_ext(v);
        for (auto& v : os.ctl_zs_last) ch.observe_ext(efrom(v));
        const FriChallenges fc = fri_challenges(ch, p.fri, degree_bits);
        // validate_proof_shape
        const size_t num_zs = nperm[i] + ctl_vars[i].size();
        const int qdf = t.quotient_degree_factor();
        if (os.local_values.size() != (size_t)t.columns || os.next_values.size() != (size_t)t.columns || os.zs.size() != num_zs || os.zs_next.size() != num_zs ||
            os.ctl_zs_last.size() != ctl_vars[i].size() || os.quotient.size() != (size_t)qdf * Config::num_challenges)
            return name + ": opening set shape";
        // eval_l_0_and_l_last, then eval_vanishing_poly at zeta
        const size_t n = (size_t)1 << degree_bits;
        const F g = gl::root_of_unity((int)degree_bits);
        E zpow = zeta;
        for (uint32_t k = 0; k < degree_bits; ++k) zpow = gl::mul(zpow, zpow);
        const E zh = gl::sub(zpow, gl::make2(1, 0));
        const E l0 = gl::mul(zh, gl::inv(gl::mul(gl::sub(zeta, gl::make2(1, 0)), (F)(n % gl::P))));
        const E ll = gl::mul(zh, gl::inv(gl::mul(gl::sub(gl::mul(zeta, g), gl::make2(1, 0)), (F)(n % gl::P))));
        XConsumer yc;
        yc.alpha[0] = XE(alpha0);
        yc.alpha[1] = XE(alpha1);
        yc.z_last = XE(gl::sub(zeta, efrom(gl::inv(g))));
        yc.lagrange_first = XE(l0);
        yc.lagrange_last = XE(ll);
        std::vector<XE> lvv, nvv;
        for (auto& v : os.local_values) lvv.push_back(XE(v));
        for (auto& v : os.next_values) nvv.push_back(XE(v));
        const XRow lv{lvv.data()}, nv{nvv.data()};
        eval_table(t, lv, nv, yc);
        const XE one((F)1);
        if (nperm[i]) {  // eval_permutation_checks (permutation.rs:302-360)
            for (size_t k = 0; k < nperm[i]; ++k) yc.constraint_first_row(XE(os.zs[k]) - one);
            std::vector<std::pair<const PermutationPair*, int>> all;
            for (auto& pp : t.permutation_pairs)
                for (uint32_t c = 0; c < Config::num_challenges; ++c) all.push_back({&pp, (int)c});
            const size_t bs = (size_t)t.permutation_batch_size();
            size_t zi = 0;
            for (size_t s = 0; s < all.size(); s += bs, ++zi) {
                XE lhs = one, rhs = one;
                for (size_t k = 0; k < bs && s + k < all.size(); ++k) {
                    const Challenge c = perm_sets[k][all[s + k].second];
                    XE l((F)0), r((F)0);
                    auto& cps = all[s + k].first->column_pairs;
                    for (size_t q = cps.size(); q-- > 0;) {
                        l = l * XE(c.beta) + lv[cps[q].first];
                        r = r * XE(c.beta) + lv[cps[q].second];
                    }
                    lhs = lhs * (l + XE(c.gamma));
                    rhs = rhs * (r + XE(c.gamma));
                }
                yc.constraint(XE(os.zs_next[zi]) * rhs - XE(os.zs[zi]) * lhs);
            }
        }
        for (auto& cv : ctl_vars[i]) {  // eval_cross_table_lookup_checks (cross_table_lookup.rs:380-419)
            auto combine = [&](const XRow& row) {
                XE s((F)0);
                for (size_t q = cv.twc->columns.size(); q-- > 0;) s = s * XE(cv.ch.beta) + eval_column(cv.twc->columns[q], row);
                return s + XE(cv.ch.gamma);
            };
            const XE lf = cv.twc->has_filter ? eval_column(cv.twc->filter, lv) : one, nf = cv.twc->has_filter ? eval_column(cv.twc->filter, nv) : one;
            yc.constraint_first_row(XE(cv.local_z) - (lf * combine(lv) + one - lf));
            yc.constraint_transition(XE(cv.next_z) - XE(cv.local_z) * (nf * combine(nv) + one - nf));
        }
        for (uint32_t j = 0; j < Config::num_challenges; ++j) {
            E s = gl::make2(0, 0);
            for (int k = qdf; k-- > 0;) s = gl::add(gl::mul(s, zpow), os.quotient[(size_t)j * qdf + k]);
            if (!eeq(yc.acc[j].v, gl::mul(zh, s))) return "Mismatch between evaluation and opening of quotient polynomial in " + name;
        }
        // fri_instance (stark.rs:87-150) and verify_fri_proof
        OpeningBatch b0, b1, b2;
        b0.point = zeta;
        for (int c = 0; c < t.columns; ++c) b0.polys.push_back({0, c});
        for (size_t c = 0; c < num_zs; ++c) b0.polys.push_back({1, (int)c});
        for (int c = 0; c < 2 * qdf; ++c) b0.polys.push_back({2, c});
        b1.point = gl::mul(zeta, g);
        for (int c = 0; c < t.columns; ++c) b1.polys.push_back({0, c});
        for (size_t c = 0; c < num_zs; ++c) b1.polys.push_back({1, (int)c});
        b2.point = efrom(gl::inv(g));
        for (size_t c = nperm[i]; c < num_zs; ++c) b2.polys.push_back({1, (int)c});
        std::vector<std::vector<E>> openings(3);
        openings[0] = os.local_values;
        openings[0].insert(openings[0].end(), os.zs.begin(), os.zs.end());
        openings[0].insert(openings[0].end(), os.quotient.begin(), os.quotient.end());
        openings[1] = os.next_values;
        openings[1].insert(openings[1].end(), os.zs_next.begin(), os.zs_next.end());
        for (F v : os.ctl_zs_last) openings[2].push_back(efrom(v));
        const std::string e = verify_fri({b0, b1, b2}, {(size_t)t.columns, num_zs, (size_t)2 * qdf}, openings, fc, {&p.trace_cap, &p.zs_cap, &p.quotient_cap}, p.fri,
                                         degree_bits);
        if (!e.empty()) return name + ": " + e;
    }
    // verify_cross_table_lookups: prod(looking Z(g^-1)) == looked Z(g^-1) per challenge pair
    std::vector<size_t> cur(T, 0);
    for (auto& ctl : sys.ctls)
        for (uint32_t k = 0; k < Config::num_challenges; ++k) {
            F prod = 1;
            for (auto& lt : ctl.looking) prod = gl::mul(prod, proofs[lt.table].openings.ctl_zs_last[cur[lt.table]++]);
            if (!ctl.has_looked) continue;
            const F looked = proofs[ctl.looked.table].openings.ctl_zs_last[cur[ctl.looked.table]++];
            if (!ctl.complete) continue;  // a side's table is outside this system: nothing to compare
            if (gl::canon(prod) != gl::canon(looked)) return "Cross-table lookup verification failed.";
        }
    return "";
}

}  // namespace verify
}  // namespace stark
}  // namespace ola
