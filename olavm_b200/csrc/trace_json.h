// Trace JSON ingest: the file `ola run` writes and `ola prove` reads
// (client/src/main.rs:166-169 serde_json::to_writer(&program.trace); :172-181 serde_json::from_reader::<Trace>) parsed into the
// flat executor records the ola_generate_* entry points take.  Host code, no CUDA.
//
// The layout is what serde derives for core/src/trace/trace.rs:320-342 `Trace`: structs are objects keyed by field name,
// fixed arrays and Vecs are arrays, GoldilocksField is a bare u64 (goldilocks_field.rs:24-26: a transparent newtype), bool is
// true / false, HashMap<String, _> is an object.  One pass over the text, no DOM: each record struct has a small schema (key ->
// slot of the flat record) and unknown keys are skipped, so the parser tolerates added fields and any key order.  Fields the
// generators recompute are skipped (PoseidonRow's round states: generation/poseidon.rs copies them, ola_generate_poseidon_trace
// replays the permutation; RangeCheckRow's limbs: the low / high 16 bits of val, trace.rs:414-418).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ola {
namespace tracejson {

struct Field {
    const char* key;
    int slot;          // first u64 of the record this key fills
    int len;           // 0: one scalar (number or bool); > 0: an array of `len` scalars
    const Field* sub;  // != nullptr: a nested object with its own schema (slots are absolute)
    int nsub;
};

// record kinds (= OLA_REC_* of include/ola_gpu.h) and their widths in u64
enum { REC_STEP = 0, REC_MEMORY, REC_RC_VAL, REC_RC_KIND, REC_BW_TAG, REC_BW_OP0, REC_BW_OP1, REC_BW_RES, REC_CMP, REC_PSDN_INPUT, REC_PSDN_FILTER,
       REC_PCHUNK, REC_STORAGE, REC_TAPE, REC_SCCALL, REC_PROG_ROW, REC_KINDS };
static constexpr uint32_t kRecWidth[REC_KINDS] = {66, 15, 1, 1, 1, 1, 1, 1, 6, 12, 4, 32, 38, 5, 24, 6};

struct Records {
    // own[k]: the records the parser produced; view k is what the generators read -- own[k], or memory borrowed from the caller
    // (ola_trace_set_records: a Rust host hands over its flattened Vec<Row> without a copy)
    std::vector<uint64_t> own[REC_KINDS];
    const uint64_t* ptr[REC_KINDS] = {};
    size_t count[REC_KINDS] = {};       // records (not u64) per kind
    size_t n_storage_access = 0;        // REC_STORAGE: builtin_storage_hash records first, builtin_program_hash after them
    uint64_t roots[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // start_end_roots

    void bind_owned() {
        for (int k = 0; k < REC_KINDS; ++k) ptr[k] = own[k].empty() ? nullptr : own[k].data(), count[k] = own[k].size() / kRecWidth[k];
    }
    const uint64_t* rows(int k) const { return ptr[k]; }
    size_t n(int k) const { return count[k]; }
    size_t u64s(int k) const { return count[k] * kRecWidth[k]; }
};

class Parser {
public:
    Parser(const char* s, size_t n) : p_(s), end_(s + n), begin_(s) {}

    void parse_trace(Records& r) {
        expect('{');
        if (peek() == '}') { ++p_; return; }
        for (;;) {
            const std::string key = parse_string();
            expect(':');
            top_level(key, r);
            if (!more('}')) break;
        }
        ws();
        if (p_ != end_) fail("trailing characters after the Trace object");
    }

private:
    const char* p_;
    const char* end_;
    const char* begin_;

    [[noreturn]] void fail(const char* what) const {
        throw std::runtime_error(std::string("trace JSON: ") + what + " at byte " + std::to_string((size_t)(p_ - begin_)));
    }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    char peek() {
        ws();
        if (p_ >= end_) fail("unexpected end of input");
        return *p_;
    }
    void expect(char c) {
        if (peek() != c) fail(c == '{' ? "expected '{'" : c == '[' ? "expected '['" : c == ':' ? "expected ':'" : "unexpected character");
        ++p_;
    }
    // after an element: ',' -> true (another follows), the closing bracket -> false
    bool more(char close) {
        const char c = peek();
        ++p_;
        if (c == ',') return true;
        if (c != close) fail("expected ',' or a closing bracket");
        return false;
    }
    std::string parse_string() {
        expect('"');
        std::string out;
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') {
                if (++p_ >= end_) fail("unterminated escape");
                switch (*p_) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'u':
                        if (end_ - p_ < 5) fail("short \\u escape");
                        out.push_back('?');  // no key or address of a Trace carries one; instruction text is skipped anyway
                        p_ += 4;
                        break;
                    default: out.push_back(*p_);
                }
                ++p_;
            } else {
                out.push_back(*p_++);
            }
        }
        if (p_ >= end_) fail("unterminated string");
        ++p_;
        return out;
    }
    void skip_string() {
        expect('"');
        while (p_ < end_ && *p_ != '"') p_ += (*p_ == '\\') ? 2 : 1;
        if (p_ >= end_) fail("unterminated string");
        ++p_;
    }
    void skip_value(int depth = 0) {
        const char c = peek();
        if (c == '"') {
            skip_string();
        } else if (c == '{' || c == '[') {
            if (depth >= 64) fail("nesting deeper than 64 levels");  // no field of a Trace nests deeper than 4; bounds the recursion
            const char close = c == '{' ? '}' : ']';
            ++p_;
            if (peek() == close) { ++p_; return; }
            for (;;) {
                if (c == '{') {
                    skip_string();
                    expect(':');
                }
                skip_value(depth + 1);
                if (!more(close)) break;
            }
        } else {
            while (p_ < end_ && *p_ != ',' && *p_ != '}' && *p_ != ']' && *p_ != ' ' && *p_ != '\n' && *p_ != '\t' && *p_ != '\r') ++p_;
        }
    }
    // a u64 (serde writes GoldilocksField, u32, u64 and u8 as bare integers) or a bool
    uint64_t parse_scalar() {
        const char c = peek();
        if (c == 't') { lit("true"); return 1; }
        if (c == 'f') { lit("false"); return 0; }
        if (c < '0' || c > '9') fail("expected an unsigned integer or a bool");
        uint64_t v = 0;
        int digits = 0;
        while (p_ < end_ && *p_ >= '0' && *p_ <= '9') {
            const uint64_t d = (uint64_t)(*p_ - '0');
            if (v > (UINT64_MAX - d) / 10) fail("integer does not fit 64 bits");
            v = v * 10 + d;
            ++p_, ++digits;
        }
        if (p_ < end_ && (*p_ == '.' || *p_ == 'e' || *p_ == 'E')) fail("expected an integer, found a float");
        return v;
    }
    void lit(const char* w) {
        const size_t n = strlen(w);
        if ((size_t)(end_ - p_) < n || memcmp(p_, w, n) != 0) fail("bad literal");
        p_ += n;
    }
    void parse_scalars(uint64_t* out, int len) {
        expect('[');
        for (int i = 0; i < len; ++i) {
            out[i] = parse_scalar();
            const bool m = more(']');
            if (m != (i + 1 < len)) fail("array of unexpected length");
        }
        if (len == 0 && peek() == ']') ++p_;
    }
    void parse_record(const Field* f, int nf, uint64_t* row) {
        expect('{');
        if (peek() == '}') { ++p_; return; }
        int next = 0;
        for (;;) {
            const char* k0;
            size_t klen;
            raw_key(k0, klen);
            expect(':');
            // serde writes the fields in declaration order, so the search resumes behind the previous hit: one comparison per
            // field for a file in that order, a full scan only for unknown keys
            const Field* hit = nullptr;
            for (int tries = 0, i = next; tries < nf; ++tries, i = i + 1 == nf ? 0 : i + 1)
                if (f[i].key[0] == k0[0] && strncmp(f[i].key, k0, klen) == 0 && f[i].key[klen] == 0) {
                    hit = &f[i];
                    next = i + 1 == nf ? 0 : i + 1;
                    break;
                }
            if (!hit)
                skip_value();
            else if (hit->sub)
                parse_record(hit->sub, hit->nsub, row);
            else if (hit->len)
                parse_scalars(row + hit->slot, hit->len);
            else
                row[hit->slot] = parse_scalar();
            if (!more('}')) break;
        }
    }
    // keys of the record structs are plain identifiers: no escapes to undo
    void raw_key(const char*& k0, size_t& klen) {
        expect('"');
        k0 = p_;
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') fail("escaped characters in a field name");
            ++p_;
        }
        if (p_ >= end_) fail("unterminated string");
        klen = (size_t)(p_ - k0);
        ++p_;
    }
    size_t parse_records(const Field* f, int nf, int rec, std::vector<uint64_t>& out) {
        expect('[');
        size_t k = 0;
        if (peek() == ']') { ++p_; return 0; }
        for (;;) {
            out.resize(out.size() + (size_t)rec, 0);
            parse_record(f, nf, out.data() + out.size() - (size_t)rec);
            ++k;
            if (!more(']')) break;
        }
        return k;
    }

    void top_level(const std::string& key, Records& r);
};

// ---- schemas: field name -> slot of the flat record (layouts documented in include/ola_gpu.h) -----------------------------------------
static const Field kRegisterSelector[] = {{"op0", 29, 0, nullptr, 0},  {"op1", 30, 0, nullptr, 0},          {"dst", 31, 0, nullptr, 0},
                                          {"aux0", 32, 0, nullptr, 0}, {"aux1", 33, 0, nullptr, 0},         {"op0_reg_sel", 35, 10, nullptr, 0},
                                          {"op1_reg_sel", 45, 10, nullptr, 0}, {"dst_reg_sel", 55, 10, nullptr, 0}};
static const Field kStep[] = {{"env_idx", 0, 0, nullptr, 0},
    {"call_sc_cnt", 1, 0, nullptr, 0},
    {"clk", 11, 0, nullptr, 0},
    {"pc", 12, 0, nullptr, 0},
    {"tp", 10, 0, nullptr, 0},
    {"addr_storage", 2, 4, nullptr, 0},
    {"addr_code", 6, 4, nullptr, 0},
    {"instruction", 25, 0, nullptr, 0},
    {"immediate_data", 28, 0, nullptr, 0},
    {"opcode", 27, 0, nullptr, 0},
    {"op1_imm", 26, 0, nullptr, 0},
    {"regs", 15, 10, nullptr, 0},
    {"register_selector", 0, 0, kRegisterSelector, 8},
    {"is_ext_line", 13, 0, nullptr, 0},
    {"ext_cnt", 14, 0, nullptr, 0},
    {"filter_tape_looking", 65, 0, nullptr, 0},
    {"storage_access_idx", 34, 0, nullptr, 0}};
static const Field kMemory[] = {{"env_idx", 0, 0, nullptr, 0},
    {"addr", 2, 0, nullptr, 0},
    {"clk", 3, 0, nullptr, 0},
    {"is_rw", 1, 0, nullptr, 0},
    {"op", 4, 0, nullptr, 0},
    {"is_write", 5, 0, nullptr, 0},
    {"diff_addr", 7, 0, nullptr, 0},
    {"diff_addr_inv", 8, 0, nullptr, 0},
    {"diff_clk", 9, 0, nullptr, 0},
    {"diff_addr_cond", 10, 0, nullptr, 0},
    {"rw_addr_unchanged", 11, 0, nullptr, 0},
    {"region_prophet", 12, 0, nullptr, 0},
    {"region_heap", 13, 0, nullptr, 0},
    {"value", 6, 0, nullptr, 0},
    {"rc_value", 14, 0, nullptr, 0}};
static const Field kRangeCheck[] = {{"val", 0, 0, nullptr, 0},
    {"filter_looked_for_mem_sort", 2, 0, nullptr, 0},
    {"filter_looked_for_mem_region", 3, 0, nullptr, 0},
    {"filter_looked_for_cpu", 1, 0, nullptr, 0},
    {"filter_looked_for_comparison", 4, 0, nullptr, 0}};
static const Field kBitwise[] = {{"opcode", 0, 0, nullptr, 0}, {"op0", 1, 0, nullptr, 0}, {"op1", 2, 0, nullptr, 0}, {"res", 3, 0, nullptr, 0}};
static const Field kCmp[] = {{"op0", 0, 0, nullptr, 0},      {"op1", 1, 0, nullptr, 0},          {"gte", 2, 0, nullptr, 0},
                             {"abs_diff", 3, 0, nullptr, 0}, {"abs_diff_inv", 4, 0, nullptr, 0}, {"filter_looking_rc", 5, 0, nullptr, 0}};
static const Field kPoseidon[] = {{"input", 0, 12, nullptr, 0},
                                  {"filter_looked_normal", 12, 0, nullptr, 0},
                                  {"filter_looked_treekey", 13, 0, nullptr, 0},
                                  {"filter_looked_storage", 14, 0, nullptr, 0},
                                  {"filter_looked_storage_branch", 15, 0, nullptr, 0}};
static const Field kPoseidonChunk[] = {{"env_idx", 0, 0, nullptr, 0}, {"clk", 1, 0, nullptr, 0},      {"opcode", 2, 0, nullptr, 0}, {"dst", 3, 0, nullptr, 0},
                                       {"op0", 4, 0, nullptr, 0},     {"op1", 5, 0, nullptr, 0},      {"acc_cnt", 6, 0, nullptr, 0}, {"value", 7, 8, nullptr, 0},
                                       {"cap", 15, 4, nullptr, 0},    {"hash", 19, 12, nullptr, 0},   {"is_ext_line", 31, 0, nullptr, 0}};
static const Field kStorageHash[] = {{"storage_access_idx", 0, 0, nullptr, 0}, {"pre_root", 1, 4, nullptr, 0}, {"root", 5, 4, nullptr, 0},
                                     {"is_write", 9, 0, nullptr, 0},           {"layer", 10, 0, nullptr, 0},   {"layer_bit", 11, 0, nullptr, 0},
                                     {"addr_acc", 12, 0, nullptr, 0},          {"addr", 13, 4, nullptr, 0},    {"pre_path", 17, 4, nullptr, 0},
                                     {"path", 21, 4, nullptr, 0},              {"hash_type", 25, 0, nullptr, 0}, {"pre_hash", 26, 4, nullptr, 0},
                                     {"hash", 30, 4, nullptr, 0},              {"sibling", 34, 4, nullptr, 0}};
static const Field kTape[] = {{"is_init", 0, 0, nullptr, 0}, {"opcode", 1, 0, nullptr, 0}, {"addr", 2, 0, nullptr, 0}, {"value", 3, 0, nullptr, 0},
                              {"filter_looked", 4, 0, nullptr, 0}};
static const Field kSCCall[] = {{"caller_env_idx", 0, 0, nullptr, 0},   {"addr_storage", 1, 4, nullptr, 0},    {"addr_code", 5, 4, nullptr, 0},
                                {"caller_op1_imm", 9, 0, nullptr, 0},   {"clk_caller_call", 10, 0, nullptr, 0}, {"clk_caller_ret", 11, 0, nullptr, 0},
                                {"regs", 12, 10, nullptr, 0},           {"callee_env_idx", 22, 0, nullptr, 0}, {"clk_callee_end", 23, 0, nullptr, 0}};
#define OLA_NF(a) ((int)(sizeof(a) / sizeof((a)[0])))

inline int hex_nibble(char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }

inline void Parser::top_level(const std::string& key, Records& r) {
    if (key == "exec") {
        parse_records(kStep, OLA_NF(kStep), 66, r.own[REC_STEP]);
    } else if (key == "memory") {
        parse_records(kMemory, OLA_NF(kMemory), 15, r.own[REC_MEMORY]);
    } else if (key == "builtin_rangecheck") {
        std::vector<uint64_t> rows;
        const size_t k = parse_records(kRangeCheck, OLA_NF(kRangeCheck), 5, rows);
        r.own[REC_RC_VAL].resize(k), r.own[REC_RC_KIND].resize(k);
        for (size_t i = 0; i < k; ++i) {
            const uint64_t* c = rows.data() + i * 5;
            r.own[REC_RC_VAL][i] = c[0];
            r.own[REC_RC_KIND][i] = c[1] ? 0 : c[2] ? 1 : c[3] ? 2 : c[4] ? 3 : 4;
        }
    } else if (key == "builtin_bitwise_combined") {
        std::vector<uint64_t> rows;
        const size_t k = parse_records(kBitwise, OLA_NF(kBitwise), 4, rows);
        r.own[REC_BW_TAG].resize(k), r.own[REC_BW_OP0].resize(k), r.own[REC_BW_OP1].resize(k), r.own[REC_BW_RES].resize(k);
        for (size_t i = 0; i < k; ++i) r.own[REC_BW_TAG][i] = rows[i * 4], r.own[REC_BW_OP0][i] = rows[i * 4 + 1], r.own[REC_BW_OP1][i] = rows[i * 4 + 2], r.own[REC_BW_RES][i] = rows[i * 4 + 3];
    } else if (key == "builtin_cmp") {
        parse_records(kCmp, OLA_NF(kCmp), 6, r.own[REC_CMP]);
    } else if (key == "builtin_poseidon") {
        std::vector<uint64_t> rows;
        const size_t k = parse_records(kPoseidon, OLA_NF(kPoseidon), 16, rows);
        r.own[REC_PSDN_INPUT].resize(k * 12), r.own[REC_PSDN_FILTER].resize(k * 4);
        for (size_t i = 0; i < k; ++i) {
            memcpy(r.own[REC_PSDN_INPUT].data() + i * 12, rows.data() + i * 16, 12 * sizeof(uint64_t));
            memcpy(r.own[REC_PSDN_FILTER].data() + i * 4, rows.data() + i * 16 + 12, 4 * sizeof(uint64_t));
        }
    } else if (key == "builtin_poseidon_chunk") {
        parse_records(kPoseidonChunk, OLA_NF(kPoseidonChunk), 32, r.own[REC_PCHUNK]);
    } else if (key == "builtin_storage_hash" || key == "builtin_program_hash") {
        // generate_storage_access_trace chains accesses then program-hash reads (storage.rs:23): keep that order whatever the file's
        std::vector<uint64_t> rows;
        const size_t k = parse_records(kStorageHash, OLA_NF(kStorageHash), 38, rows);
        if (key == "builtin_storage_hash") {
            r.own[REC_STORAGE].insert(r.own[REC_STORAGE].begin(), rows.begin(), rows.end());
            r.n_storage_access = k;
        } else {
            r.own[REC_STORAGE].insert(r.own[REC_STORAGE].end(), rows.begin(), rows.end());
        }
    } else if (key == "tape") {
        parse_records(kTape, OLA_NF(kTape), 5, r.own[REC_TAPE]);
    } else if (key == "sc_call") {
        parse_records(kSCCall, OLA_NF(kSCCall), 24, r.own[REC_SCCALL]);
    } else if (key == "start_end_roots") {
        expect('[');
        parse_scalars(r.roots, 4);
        if (!more(']')) fail("start_end_roots is a pair");
        parse_scalars(r.roots + 4, 4);
        if (more(']')) fail("start_end_roots is a pair");
    } else if (key == "addr_program_hash") {
        // HashMap<String, Vec<GoldilocksField>>: hex(4 x 8 big-endian bytes) -> the program's words (decode_addr,
        // core/src/types/merkle_tree/mod.rs:150-189).  The Rust iterates its HashMap in an unspecified order; here: file order.
        expect('{');
        if (peek() == '}') { ++p_; return; }
        for (;;) {
            const std::string hex = parse_string();
            if (hex.size() != 64) fail("a program address is 64 hex digits");
            uint64_t addr[4] = {0, 0, 0, 0};
            for (int i = 0; i < 64; ++i) {
                const int nib = hex_nibble(hex[(size_t)i]);
                if (nib < 0) fail("a program address is 64 hex digits");
                addr[i / 16] = (addr[i / 16] << 4) | (uint64_t)nib;
            }
            expect(':');
            expect('[');
            uint64_t pc = 0;
            if (peek() == ']') {
                ++p_;
            } else {
                for (;;) {
                    const uint64_t w = parse_scalar();
                    r.own[REC_PROG_ROW].insert(r.own[REC_PROG_ROW].end(), {addr[0], addr[1], addr[2], addr[3], pc, w});
                    ++pc;
                    if (!more(']')) break;
                }
            }
            if (!more('}')) break;
        }
    } else {
        skip_value();  // instructions, raw_instructions, raw_binary_instructions, builtin_storage, ret: not read by generate_traces
    }
}

inline void parse(const char* json, size_t len, Records& out) {
    Parser(json, len).parse_trace(out);
    out.bind_owned();
}

}  // namespace tracejson
}  // namespace ola
