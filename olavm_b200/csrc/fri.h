// FRI prover entry (internal).
#pragma once
#include <memory>
#include <vector>

#include "batch.h"
#include "common.h"
#include "stark_types.h"

namespace ola {
namespace fri {

// FriInstanceInfo (plonky2/plonky2/src/fri/structure.rs): opening batches of (point, [(oracle, polynomial)])
struct BatchInfo {
    stark::E point;
    std::vector<std::pair<int, int>> polys;
};
struct Instance {
    std::vector<BatchInfo> batches;
};

// PolynomialBatch::prove_openings (fri/oracle.rs:167-241) + fri_proof (fri/prover.rs:20-70)
stark::FriProof prove_openings(ola_ctx* ctx, const Instance& inst, const std::vector<const ola_batch*>& oracles, stark::Challenger& ch,
                               uint32_t degree_bits);

}  // namespace fri
}  // namespace ola
