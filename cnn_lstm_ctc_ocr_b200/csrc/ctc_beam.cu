// placeholder until the beam-search kernel lands (next commit)
#include "common.cuh"
extern "C" int ocr_ctc_beam_search_workspace_bytes(int, int, int, int, size_t* bytes) { if (bytes) *bytes = 0; return OCR_OK; }
extern "C" int ocr_ctc_beam_search(const float*, int, int, int, const int32_t*, int, int, int, int, int64_t*, int32_t*, float*, void*, size_t, ocr_stream_t)
{ ocr::set_error("ocr_ctc_beam_search: not implemented yet"); return OCR_EINVAL; }
