// STARK prover entry (internal).
#pragma once
#include <vector>

#include "common.h"
#include "stark_types.h"

namespace ola {
namespace stark {
// prove_with_traces + write_all_proof; traces[i] column-major [columns_i][2^log_ns[i]] (host or device);
// compress_challenges empty or one per table (used by Bitwise / Program)
std::vector<uint8_t> prove_all(ola_ctx* ctx, const std::vector<int>& table_ids, const std::vector<const uint64_t*>& traces, bool on_device,
                               const std::vector<uint32_t>& log_ns, const std::vector<uint64_t>& compress_challenges, const Config& cfg,
                               TranscriptHost* transcript_host = nullptr);
}  // namespace stark
}  // namespace ola
