/* ORACLE (test infrastructure, NOT product code) -- C entry points of the restated STARK prover/verifier. */
#include <cstdio>

#include "tables.hpp"

using namespace orc;

static void set_err(char* err, size_t cap, const std::string& e) {
    if (err && cap) snprintf(err, cap, "%s", e.c_str());
}

extern "C" {

/* prove_with_traces + Buffer::write_all_proof.  table_ids: ntables ids in enum order; traces[i]: column-major
 * [columns_i][2^log_ns[i]]; compress_challenges: NULL or one per table (Bitwise / Program entries used).  Returns 0 and the proof bytes, or -1 with a message. */
int orc_stark_prove(const int* table_ids, uint32_t ntables, const uint64_t* const* traces, const uint32_t* log_ns,
                    const uint64_t* compress_challenges, int check_degree, uint8_t* out, size_t cap, size_t* out_len, char* err, size_t errcap) {
    try {
        VF cc;
        if (compress_challenges) cc.assign(compress_challenges, compress_challenges + ntables);
        System sys = make_system(std::vector<int>(table_ids, table_ids + ntables), cc);
        Config cfg;
        cfg.check_quotient_degree = check_degree != 0;
        std::vector<VF> tr(ntables);
        std::vector<size_t> ns(ntables);
        for (uint32_t i = 0; i < ntables; i++) {
            ns[i] = (size_t)1 << log_ns[i];
            tr[i].assign(traces[i], traces[i] + ns[i] * sys.tables[i].columns);
            for (auto& x : tr[i]) x = gl_canon(x);
        }
        AllProof ap;
        std::string e = prove_with_traces(sys, cfg, tr, ns, ap);
        if (!e.empty()) { set_err(err, errcap, e); return -1; }
        Writer w;
        w.all(ap);
        *out_len = w.buf.size();
        if (w.buf.size() > cap) { set_err(err, errcap, "output buffer too small"); return -2; }
        memcpy(out, w.buf.data(), w.buf.size());
        return 0;
    } catch (const std::exception& ex) {
        set_err(err, errcap, ex.what());
        return -1;
    }
}

/* Buffer::read_all_proof + verify_proof.  Returns 0 if the proof verifies. */
int orc_stark_verify(const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t len, char* err, size_t errcap) {
    try {
        Config cfg;
        Reader r(proof, len);
        AllProof ap = r.all();
        if (!r.ok || r.pos != len) { set_err(err, errcap, "malformed proof bytes"); return -1; }
        /* verifier.rs:78-86: the bitwise / program compress challenges are taken from the proof */
        System sys = make_system(std::vector<int>(table_ids, table_ids + ntables), ap.compress_challenges);
        std::string e = verify_all(sys, cfg, ap);
        if (!e.empty()) { set_err(err, errcap, e); return -1; }
        return 0;
    } catch (const std::exception& ex) {
        set_err(err, errcap, ex.what());
        return -1;
    }
}

/* The reference's per-table acceptance test ("all constraints vanish on a real trace", e.g. cpu_stark.rs:974-1105,
 * memory_stark.rs tests): evaluate the table's AIR (eval_packed_generic) on every row pair (i, i+1 mod n) over the base
 * field with the row flags a ConstraintConsumer gets there (z_last = 0 on the last row, lagrange_first / _last on the
 * first / last).  Returns 0 when every constraint vanishes; 1 and the first offending (row, position of the constraint
 * in evaluation order) otherwise; -1 on error.  CTL / permutation checks are not part of it. */
int orc_air_first_failure(int table_id, const uint64_t* trace, uint32_t log_n, uint64_t compress_challenge, uint64_t* row_out,
                          int* constraint_out) {
    try {
        VF cc(1, compress_challenge);
        System sys = make_system(std::vector<int>(1, table_id), cc);
        const Table& t = sys.tables[0];
        const size_t n = (size_t)1 << log_n;
        std::vector<P<FOps>> lv(t.columns), nv(t.columns);
        for (size_t i = 0; i < n; i++) {
            const size_t j = (i + 1) % n;
            for (int c = 0; c < t.columns; c++) {
                lv[c] = P<FOps>(gl_canon(trace[(size_t)c * n + i]));
                nv[c] = P<FOps>(gl_canon(trace[(size_t)c * n + j]));
            }
            Consumer<FOps> cons(VF(1, 1), P<FOps>::c(i + 1 == n ? 0 : 1), P<FOps>::c(i == 0 ? 1 : 0), P<FOps>::c(i + 1 == n ? 1 : 0));
            t.eval_base(lv.data(), nv.data(), cons);
            if (cons.first_nonzero >= 0) {
                *row_out = i;
                *constraint_out = cons.first_nonzero;
                return 1;
            }
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

/* The compress challenge of generate_bitwise_trace (circuits/src/generation/builtin.rs:118-131) / generate_prog_trace
 * (generation/prog.rs:23-29): Challenger::new(), observe_elements(column) column after column, get_challenge(). */
uint64_t orc_compress_challenge(const uint64_t* const* cols, uint32_t ncols, size_t n) {
    Challenger ch;
    for (uint32_t c = 0; c < ncols; c++)
        for (size_t i = 0; i < n; i++) ch.observe(gl_canon(cols[c][i]));
    return ch.get_challenge();
}

/* The individual constraint values of a table's eval_packed_generic on one (local, next) row pair over the base field, in
 * emission order, BEFORE the consumer weighs them: vals[k] = the argument of the k-th yield_constr call, kinds[k] = 0
 * constraint, 1 constraint_transition, 2 constraint_first_row, 3 constraint_last_row.  Returns the number of constraints (or
 * -1); writes at most cap of them.  Lets a test pin order AND value of every constraint of another transcription. */
int orc_air_constraints(int table_id, const uint64_t* lv_in, const uint64_t* nv_in, uint64_t compress_challenge, uint64_t* vals, int* kinds, int cap) {
    try {
        VF cc(1, compress_challenge);
        System sys = make_system(std::vector<int>(1, table_id), cc);
        const Table& t = sys.tables[0];
        std::vector<P<FOps>> lv(t.columns), nv(t.columns);
        for (int c = 0; c < t.columns; c++) {
            lv[c] = P<FOps>(gl_canon(lv_in[c]));
            nv[c] = P<FOps>(gl_canon(nv_in[c]));
        }
        /* z_last = lagrange_first = lagrange_last = 1 and alpha = 0 would lose the values: record them instead */
        Consumer<FOps> cons(VF(1, 1), P<FOps>::c(1), P<FOps>::c(1), P<FOps>::c(1));
        cons.record = true;
        t.eval_base(lv.data(), nv.data(), cons);
        const int n = (int)cons.rec_raw.size();
        for (int k = 0; k < n && k < cap; k++) {
            vals[k] = gl_canon(cons.rec_raw[k]);
            kinds[k] = cons.rec_kinds[k];
        }
        return n;
    } catch (...) {
        return -1;
    }
}

int orc_table_columns(int table_id) {
    try { return table_by_id(table_id).columns; } catch (...) { return -1; }
}

}  // extern "C"
