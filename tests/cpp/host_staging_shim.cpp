// Test-only C wrappers around the host operator's staging arithmetic (csrc/dfd_host_staging.h), compiled with plain g++
// by tests/test_host_staging.py so that the CPU suite runs the very code dfd_exec.cu runs.
#include "dfd_host_staging.h"

extern "C" {
void t_append_bits(uint8_t* dst, int64_t at, const uint8_t* src, int64_t lo, int64_t n) { dfd::host::append_bits(dst, at, src, lo, n); }
int64_t t_view_offsets(const uint8_t* views, const uint8_t* valid, int64_t lo, int64_t n, int32_t* off32) {
    return dfd::host::view_offsets(views, valid, lo, n, off32);
}
void t_view_bytes(const uint8_t* views, const void* const* data_buffers, int64_t lo, int64_t n, const int32_t* off32, char* out) {
    dfd::host::view_bytes(views, data_buffers, lo, n, off32, out);
}
void t_build_views(const int32_t* off, const uint8_t* data, int64_t rows, uint8_t* views) { dfd::host::build_views(off, data, rows, views); }
int64_t t_split_list_rows(const int32_t* loff, const int32_t* coff, const uint8_t* cvalid, int64_t cvalid_offset, int64_t lo, int64_t n, int32_t* len_off,
                          int32_t* bytes_off, int32_t* lengths, int32_t* valid_off, char* valid_bytes) {
    return dfd::host::split_list_rows(loff, coff, cvalid, cvalid_offset, lo, n, len_off, bytes_off, lengths, valid_off, valid_bytes);
}
int64_t t_split_list_rows_fixed(const int32_t* loff, int32_t w, const uint8_t* cvalid, int64_t cvalid_offset, int64_t lo, int64_t n, int32_t* len_off,
                                int32_t* bytes_off, int32_t* lengths, int32_t* valid_off, char* valid_bytes) {
    return dfd::host::split_list_rows_fixed(loff, w, cvalid, cvalid_offset, lo, n, len_off, bytes_off, lengths, valid_off, valid_bytes);
}
}
