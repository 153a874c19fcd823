// mhb_internal.h -- declarations shared by the translation units of libmhb (not part of the C ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

int mhb_set_error(int code, const char *fmt, ...);

// mhb_sort_records + optional per-pass timings (host array of n_bytes doubles, ms; forces a stream sync)
int mhb_sort_records_impl(void *stream, uint32_t *a, uint32_t *b, uint64_t n, uint32_t words, const uint8_t *bytes,
                          uint32_t n_bytes, const uint64_t *first_hist, void *ws, size_t ws_bytes, int *result_in_b,
                          double *pass_ms_host);
