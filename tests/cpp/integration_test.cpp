// The reference's integration tests (tests/integrations_tests.rs:27-149) and the README-style
// flow, restated against the C++ mirror (include/probly_b200.hpp) — runs on a GPU.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "probly_b200.hpp"

using namespace probly;

struct Doc { size_t id; std::string title, description; };

static std::vector<std::string> tokenizer(std::string_view s) {      // s.split(' ')
  std::vector<std::string> out;
  size_t a = 0;
  for (;;) {
    size_t b = s.find(' ', a);
    if (b == std::string_view::npos) { out.emplace_back(s.substr(a)); break; }
    out.emplace_back(s.substr(a, b - a));
    a = b + 1;
  }
  return out;
}
static std::vector<std::string_view> title_extract(const Doc& d) { return {d.title}; }
static std::vector<std::string_view> description_extract(const Doc& d) { return {d.description}; }

#define EXPECT(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static void test_add_query_delete_bm25() {
  Index<size_t> index(2);
  Doc doc_1{0, "abc", "dfg"}, doc_2{1, "dfgh", "abcd"};
  std::vector<FieldAccessor<Doc>> acc{title_extract, description_extract};
  index.add_document(acc, tokenizer, doc_1.id, doc_1);
  index.add_document(acc, tokenizer, doc_2.id, doc_2);
  auto calc = score::bm25::make();
  auto result = index.query("abc", calc, tokenizer, {1., 1.});
  EXPECT(result.size() == 2);
  EXPECT((result[0] == QueryResult<size_t>{0, 0.6931471805599453}));
  EXPECT((result[1] == QueryResult<size_t>{1, 0.28104699650060755}));
  index.remove_document(doc_1.id);
  index.vacuum();
  result = index.query("abc", calc, tokenizer, {1., 1.});
  EXPECT(result.size() == 1);
  EXPECT((result[0] == QueryResult<size_t>{1, 0.1166450426074421}));
}

static void test_add_query_delete_zero_to_one() {
  Index<size_t> index(2);
  Doc doc_1{0, "abc", "dfg"}, doc_2{1, "dfgh", "abcd"};
  std::vector<FieldAccessor<Doc>> acc{title_extract, description_extract};
  index.add_document(acc, tokenizer, doc_1.id, doc_1);
  index.add_document(acc, tokenizer, doc_2.id, doc_2);
  auto calc = score::zero_to_one::make();
  auto result = index.query("abc", calc, tokenizer, {1., 1.});
  EXPECT(result.size() == 2);
  EXPECT((result[0] == QueryResult<size_t>{0, 1.}));
  EXPECT((result[1] == QueryResult<size_t>{1, 0.75}));
  index.remove_document(doc_1.id);          // no vacuum: the removed-mask path
  result = index.query("abc", calc, tokenizer, {1., 1.});
  EXPECT(result.size() == 1);
  EXPECT((result[0] == QueryResult<size_t>{1, 0.75}));
}

int main() {
  try {
    test_add_query_delete_bm25();
    test_add_query_delete_zero_to_one();
  } catch (const Error& e) {
    std::fprintf(stderr, "probly::Error %d: %s\n", e.code, e.what());
    return 2;
  }
  std::puts("cpp integration tests OK");
  return 0;
}
