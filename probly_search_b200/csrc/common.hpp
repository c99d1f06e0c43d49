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

// The host builder's append log as the device-side flatten sees it (pb_index_create_from_builder): one tuple per
// (doc, distinct term) in document order, the docs' field lengths, and the DFS ordinal of every builder term id.
struct LogTuple { uint32_t term, doc, tf[PB_MAX_FIELDS]; };
struct BuilderLogView {
  uint32_t F;
  const LogTuple* tuples; uint64_t n_tuples;     // the tuples of the docs >= from_doc
  const uint32_t* doc_fl; uint64_t n_docs;       // [n_docs * F]
  const uint32_t* ord_of; uint64_t n_ord;        // builder term id -> term ordinal of this flatten (0xFFFFFFFF: absent)
};
// everything of the image but the posting columns (post_blocks = NULL, max_tf / max_fl = 0) + the log view
int builder_flatten_structure(pb_builder* b, uint64_t from_doc, pb_index_image* im, BuilderLogView* view);

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
