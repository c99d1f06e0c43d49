// Shared host-side helpers of the product library (error reporting across the C ABI).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <new>

#include "../../include/probly_b200.h"

namespace pb {

// thread-local message behind pb_last_error()
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
const char* get_error();

bool utf8_valid(const uint8_t* p, size_t n);

// structural validation of a flattened image (image_io.cpp): PB_OK or PB_ERR_INVALID + message
int validate_image(const pb_index_image* im);

}  // namespace pb

// Nothing may throw across the C ABI (include/probly_b200.h "Errors").
#define PB_TRY(...)                                                    \
  try __VA_ARGS__ catch (const std::bad_alloc&) {                           \
    pb::set_error("out of host memory");                               \
    return PB_ERR_NOMEM;                                               \
  } catch (const std::exception& e) {                                  \
    pb::set_error("internal error: %s", e.what());                     \
    return PB_ERR_INVALID;                                             \
  } catch (...) {                                                      \
    pb::set_error("internal error");                                   \
    return PB_ERR_INVALID;                                             \
  }
