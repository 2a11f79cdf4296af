/* libola_gpu -- C ABI of the B200 (sm_100a) STARK proving backend for OlaVM's `circuits` crate.
 *
 * This header is the drop-in boundary.  Plain pointers and sizes only; no
 * torch / C++ types.  Every entry point names the reference interface it replaces (paths relative to
 * the reference repository root, Sin7Y/olavm @ 5d77237).  INTEGRATION.md shows the Rust `extern "C"`
 * block and the shim bodies a maintainer adds to plonky2/plonky2/src/fri/oracle.rs and
 * circuits/src/stark/prover.rs.
 *
 * Conventions
 *   - Field elements are Goldilocks `u64` (GoldilocksField is #[repr(transparent)] over u64,
 *     plonky2/field/src/goldilocks_field.rs:24-26).  Inputs may be non-canonical (>= p); outputs are
 *     always canonical.
 *   - Polynomial batches are COLUMN-MAJOR: `cols` points at ncols contiguous columns of n = 2^log_n
 *     u64 each (== Vec<PolynomialValues<F>> flattened; circuits/src/stark/prover.rs:82).
 *   - Hashes are 4 u64 (HashOut<F>, 32 bytes); Merkle caps are 2^cap_height hashes.
 *   - Extension elements are 2 u64 (c0, c1) of F[X]/(X^2-7).
 *   - Every function returns OLA_OK (0) or a negative error code; ola_gpu_last_error() gives text.
 *     Nothing throws or aborts across the ABI.  A context is not re-entrant: one host thread per ctx
 *     (the reference serialises its GPU with a global mutex, plonky2/field/src/cfft/ntt/mod.rs:48-50).
 *   - There is NO CPU fallback: without a CUDA device ola_gpu_init fails with OLA_ERR_NO_DEVICE.
 */
#ifndef OLA_GPU_H
#define OLA_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define OLA_OK 0
#define OLA_ERR_NO_DEVICE -1      /* no CUDA device / wrong architecture */
#define OLA_ERR_CUDA -2           /* CUDA runtime error (see last_error) */
#define OLA_ERR_INVALID_ARG -3    /* size not a power of two, exceeds two-adicity 32, null pointer, ... */
#define OLA_ERR_OOM -4
#define OLA_ERR_QUOTIENT_DEGREE -5 /* "Quotient has failed, the vanishing polynomial is not divisible by Z_H"
                                      (circuits/src/stark/prover.rs:469-473 panics here) */
#define OLA_ERR_ZETA_IN_SUBGROUP -6 /* "Opening point is in the subgroup." (prover.rs:508-511) */
#define OLA_ERR_INTERNAL -7

typedef struct ola_ctx ola_ctx;     /* one per GPU; owns streams, twiddles, scratch */
typedef struct ola_batch ola_batch; /* a committed PolynomialBatch resident in HBM */

/* ---- lifecycle.  Replaces gpu_init / gpu_free (plonky2/field/src/cfft/ntt/mod.rs:21-45, :55-101),
 * called once from OlaStark::default() (circuits/src/stark/ola_stark.rs:47). ---- */
int ola_gpu_init(int device, ola_ctx** out);
void ola_gpu_destroy(ola_ctx* ctx);
const char* ola_gpu_last_error(const ola_ctx* ctx);
int ola_gpu_sync(ola_ctx* ctx);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
uint64_t ola_gpu_kernel_launches(const ola_ctx* ctx);
/* the cudaStream_t every kernel of this context is launched on (for CUDA-event timing by the caller) */
void* ola_gpu_stream(ola_ctx* ctx);

/* C::Hasher of the GenericConfig (plonky2/plonky2/src/plonk/config.rs:96-111) this context commits and proves with:
 *   OLA_HASH_POSEIDON  PoseidonGoldilocksConfig (config.rs:115-122; the default; `ola prove`, client/src/main.rs:31)
 *   OLA_HASH_BLAKE3    Blake3GoldilocksConfig   (config.rs:153-161; the reference's criterion benches,
 *                      circuits/benches/fibo_loop.rs:26, and integration tests, circuits/src/stark/ola_stark.rs:684)
 * It selects leaf hashing (H::hash_no_pad), node hashing (H::two_to_one), the challenger's permutation (H::Permutation)
 * and the hash wire format (a BytesHash<32> is its 32 bytes, 4 little-endian u64 here, never reduced) of ola_hash_rows,
 * ola_merkle_rows, ola_commit*, ola_prove.  C::InnerHasher (the FRI proof-of-work hash) is Poseidon in both.
 * Under BLAKE3 the canonical u64 of each element is hashed (hash/blake3.rs:210-213 hashes the in-memory word, which is
 * a non-canonical representative with probability ~2^-32 per element); leaves are limited to 2048 elements. */
#define OLA_HASH_POSEIDON 0
#define OLA_HASH_BLAKE3 1
int ola_set_hasher(ola_ctx* ctx, int hasher);
int ola_get_hasher(const ola_ctx* ctx);

/* Per-kernel CUDA-event tracing on the context's stream (device analogue of the reference's TimingTree,
 * plonky2/plonky2/src/util/timing.rs).  begin: start recording an event pair around every launch;
 * end: synchronise, stop recording and write {"kernel": {"ms": total, "launches": n}, ...} as JSON text. */
int ola_profile_begin(ola_ctx* ctx);
int ola_profile_end(ola_ctx* ctx, char* json_out, size_t cap);

/* ---- raw device buffers (u64 elements), for callers that keep data resident between calls ---- */
int ola_dev_alloc(ola_ctx* ctx, size_t n_u64, uint64_t** dptr);
int ola_dev_free(ola_ctx* ctx, uint64_t* dptr);
int ola_dev_upload(ola_ctx* ctx, uint64_t* dst_dev, const uint64_t* src_host, size_t n_u64);
int ola_dev_download(ola_ctx* ctx, uint64_t* dst_host, const uint64_t* src_dev, size_t n_u64);
int ola_dev_copy(ola_ctx* ctx, uint64_t* dst_dev, const uint64_t* src_dev, size_t n_u64); /* device -> device, async */
/* rows [first_row, first_row+count) of a COLUMN-major device matrix (element (r,c) at cols_dev[c*col_stride + r])
 * -> host, row-major [count][ncols].  This is how opened leaves leave the GPU (MerkleTree::get,
 * merkle_tree/mod.rs:268; fri_prover_query_round, fri/prover.rs:179-181). */
int ola_dev_gather_rows(ola_ctx* ctx, const uint64_t* cols_dev, size_t col_stride, size_t ncols, size_t first_row,
                        size_t count, uint64_t* out_host);

/* ---- Goldilocks NTT family.  `on_device` != 0: pointers are device pointers (resident path);
 * == 0: host pointers, the call stages H2D/D2H itself (the reference-facing path).
 * All are batched over ncols column-major columns and work in place unless an `out` is given. ---- */

/* cfft::evaluate_poly (plonky2/field/src/cfft/mod.rs:22, serial.rs:9-15; PolynomialCoeffs::fft,
 * polynomial/mod.rs:278-283): coefficients (natural) -> values on H (natural). */
int ola_ntt_forward(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n);
/* cfft::interpolate_poly (cfft/mod.rs:128, serial.rs:52-63; PolynomialValues::ifft, polynomial/mod.rs:60-65):
 * values on H (natural) -> coefficients (natural), scaled by 1/n. */
int ola_ntt_inverse(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n);
/* cfft::evaluate_poly_with_offset (cfft/mod.rs:65, serial.rs:20-50; coset_fft_with_options,
 * polynomial/mod.rs:302-311): coefficients (natural, n each) -> values on shift*<g_{n*2^rate_bits}>.
 * out is [ncols][n << rate_bits].  natural_order != 0: out[k] = p(shift*g^k) exactly as the reference
 * returns it; == 0: "leaf order" out[r] = p(shift*g^bitrev(r)), the order the Merkle leaves use after
 * reverse_index_bits_in_place (fri/oracle.rs:84-85) -- the resident layout of this library. */
int ola_coset_lde(ola_ctx* ctx, const uint64_t* coeffs, uint64_t* out, int on_device, size_t ncols, uint32_t log_n,
                  uint32_t rate_bits, uint64_t shift, int natural_order);
/* cfft::interpolate_poly_with_offset (cfft/mod.rs:180, serial.rs:65-79; PolynomialValues::coset_ifft,
 * polynomial/mod.rs:69-74): values on shift*H (natural) -> coefficients (natural). */
int ola_coset_intt(ola_ctx* ctx, uint64_t* data, int on_device, size_t ncols, uint32_t log_n, uint64_t shift);

/* ---- Poseidon-Goldilocks (plonky2/plonky2/src/hash/poseidon.rs:593, hashing.rs:66-108) ---- */
/* states: [nstates][12] u64, permuted in place (Poseidon::poseidon). */
int ola_poseidon_permute(ola_ctx* ctx, uint64_t* states, int on_device, size_t nstates);
/* hash_no_pad of every row of a ROW-major [nrows][ncols] matrix -> digests [nrows][4]
 * (MerkleTree::new_v2 leaf loop, merkle_tree/mod.rs:186-201). */
int ola_hash_rows(ola_ctx* ctx, const uint64_t* rows, uint64_t* digests, int on_device, size_t nrows, size_t ncols);
/* MerkleTree::new_v2 over a ROW-major leaf matrix (merkle_tree/mod.rs:180-266): writes the cap
 * (2^cap_height hashes) and, if nodes_out != NULL, all heap-ordered nodes [2*nrows][4]
 * (node 1 = root, children of i at 2i/2i+1, leaf digests at nrows..2*nrows-1; node 0 unused). */
int ola_merkle_rows(ola_ctx* ctx, const uint64_t* rows, int on_device, size_t nrows, size_t ncols, uint32_t cap_height,
                    uint64_t* cap_out_host, uint64_t* nodes_out_host);

/* PolynomialBatch::from_values / from_coeffs up to and including `lde_values` (fri/oracle.rs:45-60, :101-129), i.e. the
 * commitment without its Merkle tree: cols ([ncols][n], host or device) -> coefficients [ncols][n] and the leaf-order LDE
 * [ncols][n << rate_bits] (shift 7) in CALLER-OWNED DEVICE buffers (ola_dev_alloc).  Host input is uploaded in column
 * chunks on a second stream, each chunk's iNTT + LDE overlapping the next chunk's copy.  With on_device != 0 and
 * cols == coeffs_out_dev the values are transformed in place (PolynomialValues::ifft consumes its input). */
int ola_lde_batch(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs, uint32_t rate_bits,
                  uint64_t* coeffs_out_dev, uint64_t* lde_out_dev);

/* ---- PolynomialBatch (plonky2/plonky2/src/fri/oracle.rs:31-38) ---- */
/* PolynomialBatch::from_values (oracle.rs:45-64; is_coeffs == 0) / from_coeffs (oracle.rs:66-99; is_coeffs != 0)
 * with blinding = false: iNTT -> coset LDE (shift 7, blowup 2^rate_bits) -> Poseidon Merkle tree.
 * The batch keeps, in HBM: coefficients [ncols][n] (natural order), LDE values [ncols][n<<rate_bits]
 * (column-major, leaf order), heap-ordered Merkle nodes.  cap_out_host receives 2^cap_height hashes. */
int ola_commit(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs,
               uint32_t rate_bits, uint32_t cap_height, ola_batch** out, uint64_t* cap_out_host);
/* One rank's COSET SHARD of the same commitment (multi-GPU, one process per GPU): only LDE cosets
 * [coset_first, coset_first + coset_count) are evaluated, hashed and reduced.  coset_count is a power of two
 * dividing 2^rate_bits, coset_first a multiple of it, and 2^cap_height >= 2^rate_bits / coset_count.  Leaf block i
 * of the tree is exactly coset 7 * g^bitrev(i) * H_n (evaluate_poly_with_offset's global bit-reversal and
 * oracle.rs:84-85's leaf bit-reversal cancel), so the shard is a contiguous leaf range = whole cap subtrees: the
 * batch holds leaves [coset_first * n, (coset_first + coset_count) * n) (accessors take indices RELATIVE to that
 * range) and cap_slots_out_host receives the global cap entries
 * [coset_first, coset_first + coset_count) * 2^cap_height / 2^rate_bits.  Assembling the cap is one all-gather of
 * those entries; Merkle paths never leave the shard (they stop at the cap). */
int ola_commit_shard(ola_ctx* ctx, const uint64_t* cols, int on_device, size_t ncols, uint32_t log_n, int is_coeffs,
                     uint32_t rate_bits, uint32_t cap_height, uint32_t coset_first, uint32_t coset_count, ola_batch** out,
                     uint64_t* cap_slots_out_host);
int ola_batch_free(ola_ctx* ctx, ola_batch* b);
/* shape queries */
size_t ola_batch_ncols(const ola_batch* b);
uint32_t ola_batch_degree_log(const ola_batch* b);
uint32_t ola_batch_rate_bits(const ola_batch* b);
/* device pointers into the batch (valid until ola_batch_free) */
const uint64_t* ola_batch_coeffs_dev(const ola_batch* b); /* [ncols][n] */
const uint64_t* ola_batch_lde_dev(const ola_batch* b);    /* [ncols][n<<rate_bits], leaf order */
const uint64_t* ola_batch_nodes_dev(const ola_batch* b);  /* [2*(n<<rate_bits)][4], heap order */
/* PolynomialBatch.polynomials -> host, [ncols][n] natural-order coefficients */
int ola_batch_get_coeffs(ola_ctx* ctx, const ola_batch* b, uint64_t* out_host);
/* merkle_tree.cap -> host */
int ola_batch_get_cap(ola_ctx* ctx, const ola_batch* b, uint64_t* cap_out_host);
/* merkle_tree.leaves[leaf_index .. +count) -> host, row-major [count][ncols]
 * (== MerkleTree::get, merkle_tree/mod.rs:268; PolynomialBatch::get_lde_values, oracle.rs:132-139,
 * takes index*step bit-reversed: pass that as leaf_index). */
int ola_batch_get_leaves(ola_ctx* ctx, const ola_batch* b, size_t leaf_index, size_t count, uint64_t* out_host);
/* MerkleTree::prove (merkle_tree/mod.rs:273-308): sibling digests bottom-up, log2(nleaves)-cap_height hashes.
 * Returns the number of siblings written (>= 0) or an error (< 0). */
int ola_batch_prove_leaf(ola_ctx* ctx, const ola_batch* b, size_t leaf_index, uint64_t* siblings_out_host);

/* ---- the STARK prover (circuits/src/stark/prover.rs) ----
 * prove_with_traces (prover.rs:79-85) followed by Buffer::write_all_proof (serialization.rs:377-393), with
 * F = Goldilocks, D = 2, C = PoseidonGoldilocksConfig and StarkConfig::standard_fast_config() (config.rs:18-30).
 *   table_ids  ntables ids of the reference's `Table` enum (ola_stark.rs:104-119) in ascending order; the proof
 *              covers exactly these tables and every registered cross-table lookup among them.  Passing all 12
 *              ids is `prove_with_traces` itself (ola_table_columns() < 0: unknown table id).
 *   traces[i]  column-major trace of table i: columns_i columns of 2^log_ns[i] u64 (host or device pointers)
 *   compress_challenges  NULL, or ntables field elements: entry i is table i's compress challenge -- the beta that
 *              trace generation drew for the Bitwise and Program tables (generation/mod.rs:183-188,
 *              bitwise_stark.rs:30-38, program_stark.rs:49-58); other entries are ignored and written as 0, like
 *              AllProof.compress_challenges (prover.rs:307-320).  NULL = all zero.
 *   check_quotient_degree  != 0: fail with OLA_ERR_QUOTIENT_DEGREE where the reference panics (prover.rs:469-473);
 *              == 0: "pipeline parity" mode for synthetic traces that do not satisfy the constraints
 *   proof_out  receives the proof bytes; *proof_len their count (also set when proof_cap is too small).
 * pow_witness is the SMALLEST valid nonce (the reference's rayon find_any returns an arbitrary valid one). */
int ola_prove(ola_ctx* ctx, const int* table_ids, uint32_t ntables, const uint64_t* const* traces, int on_device,
              const uint32_t* log_ns, const uint64_t* compress_challenges, int check_quotient_degree, uint8_t* proof_out,
              size_t proof_cap, size_t* proof_len);
/* ---- the same proof with the Fiat-Shamir transcript kept by the CALLER (the staged seams) ----
 * A Rust host that wants to keep its own `Challenger` (plonky2/plonky2/src/iop/challenger.rs) -- to bind the proof to more
 * public data, to share one transcript between provers, or to audit the prover's transcript -- drives the proof as a
 * sequence of transcript events instead of handing the transcript to the library.  ola_prove_session_begin starts the prover
 * (on a worker thread of the library); the caller then loops on ola_prove_session_next:
 *   OLA_EV_OBSERVE    absorb elems[0..count) in order (challenger.observe_elements); then call next again
 *   OLA_EV_CHALLENGE  squeeze `count` elements (challenger.get_n_challenges(count)) and hand them over with
 *                     ola_prove_session_supply; then call next again
 *   OLA_EV_COMPACT    challenger.compact() (prover.rs:344); then call next again
 *   OLA_EV_DONE       collect the proof with ola_prove_session_finish
 * `stage` names the seam of prove_with_traces / prove_single_table / fri_proof the event belongs to (OLA_STAGE_*, the
 * reference line it mirrors is listed with each) and `table` the Table id being proven (-1 outside prove_single_table):
 * OLA_STAGE_ZS_CAP is the boundary SURVEY 8b calls ola_ctl_z, OLA_STAGE_ALPHAS .. QUOTIENT_CAP ola_quotient, OLA_STAGE_ZETA ..
 * OPENINGS ola_open, OLA_STAGE_FRI_* ola_fri_*.  With the reference's own Challenger on the other side the bytes equal
 * ola_prove's.  The elems pointer is valid until the next call.  No callback enters the host; finish may be called early to
 * abandon a proof.  Single-GPU contexts only; the context must not be used by other calls while a session is open. */
typedef struct ola_session ola_session;
typedef struct {
    int kind;              /* OLA_EV_* */
    int stage;             /* OLA_STAGE_* */
    int table;             /* Table id, or -1 */
    const uint64_t* elems; /* OLA_EV_OBSERVE: canonical field elements to absorb */
    size_t count;          /* OBSERVE: number of elems; CHALLENGE: number of elements to supply */
} ola_transcript_event;
#define OLA_EV_OBSERVE 1
#define OLA_EV_CHALLENGE 2
#define OLA_EV_COMPACT 3
#define OLA_EV_DONE 4
#define OLA_EV_FAILED 5
#define OLA_STAGE_TRACE_CAPS 1        /* prover.rs:147-150 */
#define OLA_STAGE_CTL_CHALLENGES 2    /* prover.rs:152-158, get_grand_product_challenge_set */
#define OLA_STAGE_TABLE_BEGIN 3       /* prover.rs:344-372: compact, permutation challenges */
#define OLA_STAGE_ZS_CAP 4            /* prover.rs:411-413 */
#define OLA_STAGE_ALPHAS 5            /* prover.rs:415 */
#define OLA_STAGE_QUOTIENT_CAP 6      /* prover.rs:489-491 */
#define OLA_STAGE_ZETA 7              /* prover.rs:493 */
#define OLA_STAGE_OPENINGS 8          /* prover.rs:530, proof.rs:248-283 */
#define OLA_STAGE_FRI_ALPHA 9         /* fri/oracle.rs:176 */
#define OLA_STAGE_FRI_LAYER_CAP 10    /* fri/prover.rs:94 */
#define OLA_STAGE_FRI_BETA 11         /* fri/prover.rs:96 */
#define OLA_STAGE_FRI_FINAL_POLY 12   /* fri/prover.rs:118 */
#define OLA_STAGE_FRI_POW 13          /* fri/prover.rs:131: get_hash */
#define OLA_STAGE_FRI_QUERY_INDICES 14 /* fri/prover.rs:157-160 */
int ola_prove_session_begin(ola_ctx* ctx, const int* table_ids, uint32_t ntables, const uint64_t* const* traces, int on_device,
                            const uint32_t* log_ns, const uint64_t* compress_challenges, int check_quotient_degree, ola_session** out);
int ola_prove_session_next(ola_session* s, ola_transcript_event* ev);
int ola_prove_session_supply(ola_session* s, const uint64_t* challenges, size_t count);
int ola_prove_session_finish(ola_session* s, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* ---- multi-GPU: one process (or thread) per GPU, the proof coset-sharded across the ranks ----
 * After ola_set_comm, ola_prove called COLLECTIVELY by every rank with the SAME arguments shards the three dominant
 * costs by LDE cosets -- commitments (coset LDE + leaf hashing + subtree reduction), constraint-quotient evaluation
 * ("next row" stays inside a coset: no communication) -- and replicates the rest (Z columns, openings, FRI layers of
 * the 2-column composition polynomial, transcript).  Exchanges per table: cap entries (all-gather of 16 digests),
 * quotient values (all-gather of 2 x 8n u64) and the opened query rows / Merkle paths, answered by the owner of each
 * leaf (all-reduce over zero-filled buffers).  Every rank returns the same, single-GPU-identical proof bytes.
 * world must divide 8 (the blowup).  The library does not link a communication library: the host supplies the two
 * collectives (NCCL in production -- see olavm_b200/dist.py for the torch.distributed binding).  Contract of both
 * callbacks: operate on DEVICE pointers of this rank's GPU, ordered after the work already enqueued on `stream`, and
 * make the result visible to work enqueued on `stream` after they return; return 0 on success. */
typedef int (*ola_allgather_fn)(void* user, const void* send_dev, void* recv_dev, size_t bytes_per_rank, void* stream);
typedef int (*ola_allreduce_u64_fn)(void* user, void* buf_dev, size_t count_u64, void* stream); /* in place, wrapping sum */
int ola_set_comm(ola_ctx* ctx, int rank, int world, ola_allgather_fn allgather, ola_allreduce_u64_fn allreduce_sum, void* user);

/* The same communicator bound natively: the library dlopens libnccl (libnccl_path, or NULL = the copy already loaded in
 * the process / the default search path) and issues ncclAllGather / ncclAllReduce on the context's stream itself, so
 * that no host-language callback sits on the proving path.  Rank 0 calls ola_nccl_unique_id and hands the 128 bytes to
 * the other ranks over any side channel; then every rank calls ola_set_comm_nccl (collective: ncclCommInitRank). */
int ola_nccl_unique_id(const char* libnccl_path, uint8_t id_out[128]);
int ola_set_comm_nccl(ola_ctx* ctx, const char* libnccl_path, int rank, int world, const uint8_t id[128]);
/* bytes this rank has received through its communicator so far (all-gathers: world * bytes_per_rank; all-reduces: 8 * count) */
uint64_t ola_comm_bytes(const ola_ctx* ctx);

/* verify_proof (circuits/src/stark/verifier.rs:32-212) over Buffer::read_all_proof's bytes: host code, no GPU or
 * context needed.  Like the reference's, it is fixed at the full system: table_ids must be the 12 ids 0..11 in order, and
 * every cross-table lookup is checked.  Returns OLA_OK when the proof is accepted; OLA_ERR_INVALID_ARG with the reason in err
 * (NUL-terminated, truncated to errcap) when it is rejected. */
int ola_verify(const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err, size_t errcap);
/* the same for a proof made under `hasher` (OLA_HASH_*): verify_proof::<F, C, D> with C = Blake3GoldilocksConfig */
int ola_verify_cfg(int hasher, const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err,
                   size_t errcap);
/* WEAKER, for tests and for proofs of a subsystem made by ola_prove with fewer tables: verifies an ordered subset of the
 * tables.  A cross-table lookup with a side outside the subset cannot be balanced and is NOT checked (its Z columns are
 * still checked against the table's own constraints), so acceptance says nothing about those lookups. */
int ola_verify_subsystem_cfg(int hasher, const int* table_ids, uint32_t ntables, const uint8_t* proof, size_t proof_len, char* err,
                             size_t errcap);
/* ---- trace-generation tail: what sits directly in front of prove_with_traces ----
 * generate_poseidon_trace (circuits/src/generation/poseidon.rs:5-130) together with the per-round states the executor
 * records for each hash (core/src/util/poseidon_utils.rs:289-420): `inputs` [nrows][12] permutation inputs and `filters`
 * [nrows][4] (looked_normal, looked_treekey, looked_storage_leaf, looked_storage_branch; NULL = all 0) -> the
 * column-major Poseidon table `out` [134][2^log_n] (builtins/poseidon/columns.rs:6-42); rows past nrows are the
 * zero-input padding row (POSEIDON_ZERO_HASH_*).  One GPU thread per row.  on_device: all three pointers are device
 * pointers. */
int ola_generate_poseidon_trace(ola_ctx* ctx, const uint64_t* inputs, const uint64_t* filters, size_t nrows, uint32_t log_n, uint64_t* out,
                                int on_device);
/* permuted_cols (circuits/src/stark/lookup.rs:68-131): the permuted input / permuted table pair of the Halo2-style lookup
 * argument (eval_lookups, lookup.rs:13-35) for one input column and one table column of n elements, n a power of two in
 * [2, 2^24].  permuted_inputs = the sorted canonical inputs; permuted_table = the reference's merge walk (matching table
 * value on the first occurrence of an input value, otherwise the most recently skipped table value, unfilled rows taking the
 * values never used, in order) -- computed here as sorts, prefix sums and a bracket matching on the GPU, same columns.
 * on_device: all four pointers are device pointers. */
int ola_permuted_cols(ola_ctx* ctx, const uint64_t* inputs, const uint64_t* table, size_t n, uint64_t* permuted_inputs, uint64_t* permuted_table,
                      int on_device);
/* generate_rc_trace (circuits/src/generation/builtin.rs:249-316) from the executor's RangeCheckRow list
 * (core/src/trace/trace.rs:401-425): vals[nrows] and kinds[nrows] (which table looks the value up: 0 cpu, 1 memory sort,
 * 2 memory region, 3 comparison -- the filter column that is 1 in that row) -> the column-major RangeCheck table
 * out [12][2^log_n] (builtins/rangecheck/columns.rs:27-40: 4 filters, val, limb_lo, limb_hi, the two permuted limb columns,
 * the fixed 0..2^16-1 column padded with its last value, its two permuted copies); log_n >= 16 and nrows <= 2^log_n.
 * on_device: all three pointers are device pointers. */
int ola_generate_rangecheck_trace(ola_ctx* ctx, const uint64_t* vals, const uint64_t* kinds, size_t nrows, uint32_t log_n, uint64_t* out,
                                  int on_device);
/* generate_bitwise_trace (circuits/src/generation/builtin.rs:35-206) from the executor's BitwiseCombinedRow list
 * (core/src/trace/trace.rs:368-399): tags[nrows] (c.opcode: 1 << Opcode::AND / OR / XOR), op0 / op1 / res[nrows] -> the
 * column-major Bitwise table out [59][2^log_n] (builtins/bitwise/columns.rs:23-62), log_n >= 18 (3 * 2^16 fixed rows), and the
 * table's compress challenge *beta_out (the (trace, beta) pair the reference returns; beta is ola_prove's
 * compress_challenges entry for the Bitwise table).  Row fill, the beta-compressed limb triples and the sixteen permuted_cols
 * run on the GPU; the challenge is the library's host transcript over the twelve limb columns.  The table is the one the
 * reference produces, including its handling of the fourth limb (columns 8, 12 and 16 stay zero: builtin.rs:66, :71, :76
 * write it to the first column of the next range, which is overwritten).  on_device: tags, op0, op1, res and out are device
 * pointers (beta_out is always a host pointer). */
int ola_generate_bitwise_trace(ola_ctx* ctx, const uint64_t* tags, const uint64_t* op0, const uint64_t* op1, const uint64_t* res, size_t nrows,
                               uint32_t log_n, uint64_t* out, uint64_t* beta_out, int on_device);
/* generate_cmp_trace (circuits/src/generation/builtin.rs:208-247): cells [nrows][6] = (op0, op1, gte, abs_diff, abs_diff_inv,
 * filter_looking_rc) as the executor recorded them (CmpRow) -> the column-major Cmp table out [6][2^log_n]
 * (builtins/cmp/columns.rs:16-22); padding rows are (1, 0, 1, 1, 1, 0). */
int ola_generate_cmp_trace(ola_ctx* ctx, const uint64_t* cells, size_t nrows, uint32_t log_n, uint64_t* out, int on_device);
/* generate_cpu_trace (circuits/src/generation/cpu.rs:11-218): the executor's Step list (core/src/trace/trace.rs) as one record
 * of 66 u64 per executed row ->  the column-major CPU table out [94][2^log_n] (circuits/src/cpu/columns.rs), padding rows
 * included (cpu.rs:180-208).  Record layout (field.0 of the Rust struct, in this order):
 *    0 env_idx   1 call_sc_cnt   2..5 addr_storage[4]   6..9 addr_code[4]   10 tp   11 clk   12 pc   13 is_ext_line   14 ext_cnt
 *   15..24 regs[10]   25 instruction   26 op1_imm   27 opcode   28 immediate_data
 *   29 op0   30 op1   31 dst   32 aux0   33 aux1  (register_selector)   34 storage_access_idx
 *   35..44 op0_reg_sel[10]   45..54 op1_reg_sel[10]   55..64 dst_reg_sel[10]   65 filter_tape_looking
 * One GPU thread per table row.  on_device: steps and out are device pointers. */
int ola_generate_cpu_trace(ola_ctx* ctx, const uint64_t* steps, size_t nrows, uint32_t log_n, uint64_t* out, int on_device);
/* generate_memory_trace (circuits/src/generation/memory.rs:8-155): the executor's MemoryTraceCell list (core/src/trace/trace.rs;
 * sorted by address and differenced by gen_memory_table) as one record of 15 u64 per cell -> the column-major Memory table
 * out [29][2^log_n] (circuits/src/memory/columns.rs:12-45), log_n >= 1, padding rows continuing the write-once region
 * (memory.rs:113-146).  Record layout:
 *    0 env_idx   1 is_rw   2 addr   3 clk   4 op   5 is_write   6 value   7 diff_addr   8 diff_addr_inv   9 diff_clk
 *   10 diff_addr_cond   11 rw_addr_unchanged   12 region_prophet   13 region_heap   14 rc_value
 * on_device: cells and out are device pointers. */
int ola_generate_memory_trace(ola_ctx* ctx, const uint64_t* cells, size_t ncells, uint32_t log_n, uint64_t* out, int on_device);
/* generate_prog_trace (circuits/src/generation/prog.rs:18-157): the Step records of ola_generate_cpu_trace (executed side: one
 * row per fetched instruction word and one per immediate, ext lines skipped), prog_rows [nprog_rows][6] = (code address 0..3, pc,
 * word) for every word of every program in the order the Rust walks `progs`, and roots[8] = start_root[4], end_root[4] (HOST
 * pointer) -> the column-major Program table out [18][2^log_n] (circuits/src/program/columns.rs:3-16) and its compress challenge
 * *beta_out (the (trace, beta) pair the reference returns).  2^log_n must hold max(fetched words, program words).  The beta-
 * compression, the placement of the executed rows (a prefix sum) and the table's permuted_cols run on the GPU.  on_device:
 * steps, prog_rows and out are device pointers (roots and beta_out stay on the host). */
int ola_generate_program_trace(ola_ctx* ctx, const uint64_t* steps, size_t nsteps, const uint64_t* prog_rows, size_t nprog_rows, const uint64_t* roots,
                               uint32_t log_n, uint64_t* out, uint64_t* beta_out, int on_device);
/* generate_poseidon_chunk_trace (circuits/src/generation/poseidon_chunk.rs:7-88): the executor's PoseidonChunkRow list
 * (core/src/trace/trace.rs:179-192), one record of 32 u64 per line -> the column-major PoseidonChunk table out [53][2^log_n]
 * (builtins/poseidon/columns.rs:42-68); the result-line, first-padding and filter columns are derived as the Rust derives them,
 * rows past nrows are padding lines.  Record layout (struct field order):
 *    0 env_idx   1 clk   2 opcode   3 dst   4 op0   5 op1   6 acc_cnt   7..14 value[8]   15..18 cap[4]   19..30 hash[12]   31 is_ext_line
 * Every ola_generate_* below: 2^log_n >= max(2, nrows) rows; on_device: rows and out are device pointers. */
int ola_generate_poseidon_chunk_trace(ola_ctx* ctx, const uint64_t* rows, size_t nrows, uint32_t log_n, uint64_t* out, int on_device);
/* generate_storage_access_trace (circuits/src/generation/storage.rs:7-123): StorageHashRow records (core/src/trace/trace.rs:279-295),
 * 38 u64 each, the n_access storage accesses first and the n_prog_reads program-hash reads after them (the Rust chains the two
 * slices, storage.rs:23) -> the column-major StorageAccess table out [48][2^log_n] (builtins/storage/columns.rs:3-33); padding rows
 * carry the last root.  Record layout:
 *    0 storage_access_idx   1..4 pre_root   5..8 root   9 is_write   10 layer   11 layer_bit   12 addr_acc   13..16 addr
 *   17..20 pre_path   21..24 path   25 hash_type   26..29 pre_hash   30..33 hash   34..37 sibling */
int ola_generate_storage_access_trace(ola_ctx* ctx, const uint64_t* rows, size_t n_access, size_t n_prog_reads, uint32_t log_n, uint64_t* out,
                                      int on_device);
/* generate_tape_trace (circuits/src/generation/tape.rs:10-73): TapeRow records (core/src/trace/trace.rs:297-304) of 5 u64
 * (is_init, opcode, addr, value, filter_looked) -> the column-major Tape table out [6][2^log_n] (builtins/tape/columns.rs:3-9);
 * padding rows repeat the last row as an unlooked TLOAD. */
int ola_generate_tape_trace(ola_ctx* ctx, const uint64_t* rows, size_t nrows, uint32_t log_n, uint64_t* out, int on_device);
/* generate_sccall_trace (circuits/src/generation/sccall.rs:11-64): SCCallRow records (core/src/trace/trace.rs:306-317) of 24 u64
 *    0 caller_env_idx   1..4 addr_storage   5..8 addr_code   9 caller_op1_imm   10 clk_caller_call   11 clk_caller_ret
 *   12..21 regs[10]   22 callee_env_idx   23 clk_callee_end
 * -> the column-major SCCall table out [26][2^log_n] (builtins/sccall/columns.rs:4-20). */
int ola_generate_sccall_trace(ola_ctx* ctx, const uint64_t* rows, size_t nrows, uint32_t log_n, uint64_t* out, int on_device);
/* generate_prog_chunk_trace (circuits/src/generation/prog.rs:158-249): prog_rows [nprog_rows][6] = (code address 0..3, pc, word)
 * for every word of every program in the order the Rust walks `progs` (the buffer ola_generate_program_trace takes; a program
 * starts where pc == 0) -> the column-major ProgChunk table out [40][2^log_n] (program/columns.rs:47-62): lines of eight words
 * absorbed by a Poseidon sponge.  The table is the one the reference produces: the unused word slots of a program's last line
 * hold the previous line's hash (overwrite-mode sponge, prog.rs:212-216) and the sponge state is NOT reset between programs
 * (prog.rs:199 declares pre_hash once), so with more than one program the later programs' first lines carry a non-zero capacity.
 * The line placement is a prefix sum, the hash chain is sequential by construction (one GPU thread). */
int ola_generate_prog_chunk_trace(ola_ctx* ctx, const uint64_t* prog_rows, size_t nprog_rows, uint32_t log_n, uint64_t* out, int on_device);
/* ---- Trace JSON ingest and the `ola prove` flow ----
 * `ola run` writes serde_json::to_writer(&program.trace) (client/src/main.rs:166-169) and `ola prove` reads it back with
 * serde_json::from_reader::<Trace> (:172-181), calls generate_traces + prove_with_traces (circuits/src/stark/prover.rs:43-66)
 * and writes Buffer::write_all_proof's bytes (:199-206).  The entry points below are that flow for a host without serde: the
 * JSON text of a core::trace::trace::Trace (trace.rs:320-342) is parsed once into the flat executor records the
 * ola_generate_* entry points take (host code, no context needed), and ola_prove_trace generates the twelve tables in device
 * memory and proves them there.  A Rust host does not need the parser: it flattens its own Trace into the same records. */
typedef struct ola_trace ola_trace;
/* Returns OLA_ERR_INVALID_ARG with the reason (and the byte offset) in err when the text is not a serialised Trace.  Unknown
 * keys are skipped; fields the generators recompute (PoseidonRow's round states, RangeCheckRow's limbs) are not read.
 * addr_program_hash (a HashMap the Rust iterates in unspecified order) is taken in file order. */
int ola_trace_from_json(const char* json, size_t len, ola_trace** out, char* err, size_t errcap);
void ola_trace_free(ola_trace* t);
#define OLA_REC_STEP 0             /* exec                      [k][66]  (ola_generate_cpu_trace's record)            */
#define OLA_REC_MEMORY 1           /* memory                    [k][15]                                               */
#define OLA_REC_RC_VAL 2           /* builtin_rangecheck        [k]      val                                          */
#define OLA_REC_RC_KIND 3          /*                           [k]      0 cpu 1 mem sort 2 mem region 3 comparison 4 none */
#define OLA_REC_BITWISE_TAG 4      /* builtin_bitwise_combined  [k]      opcode                                       */
#define OLA_REC_BITWISE_OP0 5
#define OLA_REC_BITWISE_OP1 6
#define OLA_REC_BITWISE_RES 7
#define OLA_REC_CMP 8              /* builtin_cmp               [k][6]                                                */
#define OLA_REC_POSEIDON_INPUT 9   /* builtin_poseidon          [k][12]  input                                        */
#define OLA_REC_POSEIDON_FILTER 10 /*                           [k][4]   normal, treekey, storage, storage_branch     */
#define OLA_REC_POSEIDON_CHUNK 11  /* builtin_poseidon_chunk    [k][32]                                               */
#define OLA_REC_STORAGE_HASH 12    /* builtin_storage_hash then builtin_program_hash  [k][38]                         */
#define OLA_REC_TAPE 13            /* tape                      [k][5]                                                */
#define OLA_REC_SCCALL 14          /* sc_call                   [k][24]                                               */
#define OLA_REC_PROG_ROW 15        /* addr_program_hash         [m][6]   (code address 0..3, pc, word)                */
#define OLA_REC_ROOTS 16           /* start_end_roots           [1][8]                                                */
#define OLA_REC_STORAGE_ACCESS_COUNT 17 /* *nrows = how many OLA_REC_STORAGE_HASH records are storage accesses (rows = NULL) */
/* The same object built from records the caller already holds (a Rust host flattening its own Trace: no JSON, no copy).
 * ola_trace_new makes an empty trace; ola_trace_set_records points kind `kind` at nrows records the CALLER keeps alive until the
 * trace is freed or the kind is set again (OLA_REC_ROOTS copies its 8 values; OLA_REC_STORAGE_ACCESS_COUNT takes the count in
 * nrows, rows ignored).  Kinds that are never set are empty lists, as in a Trace::default(). */
int ola_trace_new(ola_trace** out);
int ola_trace_set_records(ola_trace* t, int kind, const uint64_t* rows, size_t nrows);
/* *rows points into the trace object (valid until ola_trace_free), *nrows records of *rec_u64 u64 each */
int ola_trace_records(const ola_trace* t, int kind, const uint64_t** rows, size_t* nrows, uint32_t* rec_u64);
/* log2 of the row count generate_traces gives table `table_id` (0..11) for this trace, or a negative error */
int ola_trace_table_log_rows(const ola_trace* t, int table_id);
/* generate_traces (circuits/src/generation/mod.rs:79-213): the twelve tables, column-major, generated in device memory.
 * tables_dev[12] receives device pointers the caller frees with ola_dev_free (table i: ola_table_columns(i) << log_ns[i] u64),
 * log_ns[12] the row counts, compress_challenges[12] the Bitwise / Program betas (other entries 0) -- the three arrays
 * ola_prove takes with on_device = 1. */
int ola_generate_traces(ola_ctx* ctx, const ola_trace* t, uint64_t** tables_dev, uint32_t* log_ns, uint64_t* compress_challenges);
/* `ola prove`: ola_generate_traces + ola_prove over the twelve tables (degree check on) + the proof bytes; the tables never
 * leave the device.  proof_out / proof_cap / proof_len as in ola_prove. */
int ola_prove_trace(ola_ctx* ctx, const ola_trace* t, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* The compress challenge of the Bitwise / Program tables: a fresh Poseidon Challenger observes `ncols` HOST columns of n
 * elements, column after column, and squeezes one element (generate_bitwise_trace, generation/builtin.rs:118-131: the 12
 * limb columns; generate_prog_trace, generation/prog.rs:23-29: the 8 interleaved root limbs as one column).  A duplex
 * sponge is sequential: host code on the library's own transcript, no context needed. */
int ola_compress_challenge(const uint64_t* const* cols, uint32_t ncols, size_t n, uint64_t* beta_out);

/* Diagnostic (host code, no context): the individual constraint values of a table's eval_packed_generic (e.g.
 * CpuStark, circuits/src/cpu/cpu_stark.rs:871-946) on ONE (local, next) row pair over the base field, in emission order and
 * before the ConstraintConsumer weighs them -- vals_out[k] is the argument of the k-th yield_constr call, kinds_out[k] = 0
 * constraint, 1 constraint_transition, 2 constraint_first_row, 3 constraint_last_row.  These are the same constraint
 * transcriptions the quotient kernels and ola_verify compile.  Returns the number of constraints (at most cap are
 * written) or a negative error. */
int ola_air_constraints(int table_id, const uint64_t* lv, const uint64_t* nv, uint64_t compress_challenge, uint64_t* vals_out, int* kinds_out,
                        int cap);
/* number of trace columns of a table (S::COLUMNS), or -1 if its constraint kernel is not compiled in */
int ola_table_columns(int table_id);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OLA_GPU_H */
