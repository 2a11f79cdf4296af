// STARK prover entry (internal).
#pragma once
#include <functional>
#include <vector>

#include "common.h"
#include "stark_types.h"

namespace ola {
namespace stark {
// A table (device-resident, on_device only) that is still being completed when prove_all starts: it is committed LAST, and just
// before that `finish` is called -- it completes the table in place on the context's stream and returns its compress challenge.
// ola_prove_trace uses it for the Bitwise table, whose challenge is a long sequential host transcript (generation/builtin.rs:118-131)
// that runs on a second host thread while the other eleven tables are committed.
struct LateTable {
    size_t index;
    std::function<uint64_t()> finish;
};
// prove_with_traces + write_all_proof; traces[i] column-major [columns_i][2^log_ns[i]] (host or device);
// compress_challenges empty or one per table (used by Bitwise / Program)
std::vector<uint8_t> prove_all(ola_ctx* ctx, const std::vector<int>& table_ids, const std::vector<const uint64_t*>& traces, bool on_device,
                               const std::vector<uint32_t>& log_ns, const std::vector<uint64_t>& compress_challenges, const Config& cfg,
                               TranscriptHost* transcript_host = nullptr, const LateTable* late = nullptr);
}  // namespace stark
}  // namespace ola
