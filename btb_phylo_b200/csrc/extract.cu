// placeholder while the extraction kernels are being written (replaced in the next commit)
#include "common.cuh"
using namespace btb;
extern "C" int64_t btb_plane_words(int64_t G) { return ((G + 31) / 32 + 3) / 4 * 4; }
extern "C" int btb_pack_planes(const uint8_t *, const int64_t *, const int32_t *, const int32_t *, int64_t, int64_t, int64_t, int64_t, uint32_t *, int32_t *, void *) { return fail(BTB_EINVAL, "not built yet"); }
extern "C" int btb_column_mask(const uint32_t *, int64_t, int64_t, int64_t, uint32_t *, int, void *) { return fail(BTB_EINVAL, "not built yet"); }
extern "C" int btb_select_columns(const uint32_t *, int64_t, int, uint32_t *, uint32_t *, uint32_t *, void *) { return fail(BTB_EINVAL, "not built yet"); }
extern "C" int btb_compact_columns(const uint32_t *, int64_t, int64_t, int64_t, int64_t, const uint32_t *, int64_t, uint8_t *, int64_t, void *) { return fail(BTB_EINVAL, "not built yet"); }
extern "C" int btb_extract_host(const uint8_t *, int64_t, int64_t, int64_t, int, int, uint8_t *, int64_t *, uint8_t *, int64_t) { return fail(BTB_EINVAL, "not built yet"); }
extern "C" int btb_snp_sites(const char *, const char *, int, int, int, int64_t *, int64_t *, int64_t *) { return fail(BTB_EINVAL, "not built yet"); }
