// placeholder, replaced below
#include "apc_internal.h"
namespace apc {
int exact_count_select(Ctx *c, uint8_t, float, uint64_t, uint64_t, const uint64_t *, uint64_t,
                       std::vector<uint64_t> &, std::vector<uint64_t> &, uint64_t *, uint64_t *) {
    return fail(c, APC_ERR_INVALID, "exact stage not built yet");
}
} // namespace apc
