// Host-side launch API of the NTT kernels (internal to libola_gpu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

struct ola_ctx;

namespace ola {
namespace ntt {

static constexpr int MAX_COSETS = 16;

// Forward (Cooley-Tukey) network over natural-order input.
//   dst[col][coset][pos]  (pos = bit-reversed position: "leaf order")   when !natural_output
//   dst[col][k]           (k natural; cosets interleaved k = kk*ncosets + bitrev(coset)) when natural_output
struct FwdDesc {
    const uint64_t* src = nullptr;  // [ncols] columns of 2^log_n elements, natural order
    uint64_t* dst = nullptr;        // final output
    uint64_t* work = nullptr;       // intermediate of multi-pass transforms (nullptr: use dst, i.e. in place);
                                    // must differ from dst when natural_output; may alias src (src is clobbered)
    size_t src_col_stride = 0, dst_col_stride = 0, dst_coset_stride = 0;
    size_t work_col_stride = 0, work_coset_stride = 0;
    size_t ncols = 0;
    int log_n = 0;
    int coset_bits = 0;        // 2^coset_bits cosets shift * g^{bitrev(i)} * H_n (blowup of an LDE)
    int coset_first = 0;       // compute only cosets [coset_first, coset_first + coset_count) (a rank's shard of the LDE);
    int coset_count = -1;      // -1: all.  Output coset k of the range lands at dst + k * dst_coset_stride.
    uint64_t shift = 1;        // domain offset
    bool inverse_roots = false;  // use omega^-1 (values -> coefficients)
    bool natural_output = false;
    bool apply_scale = false;  // multiply every output by `scale` (1/n for an inverse transform)
    uint64_t scale = 1;
    const char* tag_strided = "ntt_strided";  // profile names of this transform's launches
    const char* tag_contig = "ntt_contig";
};

void init_twiddles(ola_ctx* ctx);
void free_twiddles(ola_ctx* ctx);
std::vector<int> plan_passes(int log_n);
void forward(ola_ctx* ctx, const FwdDesc& d);
// Inverse (Gentleman-Sande) network, in place: values at bit-reversed positions on shift*H_n -> natural
// coefficients (scaled by 1/n, coset shift removed).
void inverse_from_leaf_order(ola_ctx* ctx, uint64_t* data, size_t col_stride, size_t ncols, int log_n, uint64_t shift);
// data[col][j] *= base * step^j
void scale_powers(ola_ctx* ctx, uint64_t* data, size_t col_stride, size_t ncols, size_t n, uint64_t base, uint64_t step);

}  // namespace ntt
}  // namespace ola
