// Linear-time regex matcher (Pike VM over UTF-8 code points), RE2-syntax subset, used ONLY on dictionary entries:
// `regexp_matches(col, 're', 'i')` = RE2 partial match, case-insensitive
// (core/src/main/scala/com/cardinal/utils/ast/BaseExpr.scala:485-486, 500-501).
// Supported: literals, '.', escapes \d \D \w \W \s \S \b \B \t \n \r \f \v \xHH \x{H..} and escaped punctuation,
// classes [a-z] [^...] with escapes and [:posix:] names, ^ $ \A \z, groups ( ) (?: ) (?i) (?P<n> ), alternation,
// quantifiers * + ? {n} {n,} {n,m} and their lazy forms.  Back-references / look-around are rejected, as in RE2.
#pragma once
#include <string>
#include <vector>

#include "lk_common.h"

namespace lk {

class Regex {
 public:
  // Throws lk::Error(LK_ERR_UNSUPPORTED) on syntax outside the subset.
  Regex(const std::string& pattern, bool case_insensitive);
  ~Regex();
  Regex(const Regex&) = delete;
  Regex& operator=(const Regex&) = delete;
  bool search(const char* s, size_t n) const;  // partial match
  bool search(const std::string& s) const { return search(s.data(), s.size()); }

 private:
  struct Range { uint32_t lo, hi; };
  struct Inst {
    enum Op : uint8_t { Char, Any, Class, Split, Jmp, Match, Bol, Eol, WordB, NWordB } op;
    uint32_t x = 0, y = 0;  // Char: code point; Class: class index (y = negated); Split/Jmp: targets
  };
  struct Node;
  std::vector<Inst> prog_;
  std::vector<std::vector<Range>> classes_;
  bool fold_;

  // parser state
  const std::string* pat_ = nullptr;
  size_t pos_ = 0;
  bool ci_ = false;
  int parse_alt();
  int parse_concat();
  int parse_repeat();
  int parse_atom();
  int parse_class();
  bool parse_escape_class(std::vector<Range>& out, bool& negated);
  uint32_t parse_escape_char();
  uint32_t next_cp();
  bool eof() const { return pos_ >= pat_->size(); }
  char cur() const { return (*pat_)[pos_]; }

  std::vector<Node> nodes_;
  int new_node(int kind);
  void emit(int node);
  int add_class(std::vector<Range> r, bool negated);
  bool class_match(const Inst& in, uint32_t cp) const;
  void add_thread(std::vector<uint32_t>& list, std::vector<uint32_t>& mark, uint32_t gen, uint32_t pc, bool at_start,
                  bool at_end, bool prev_word, bool next_word) const;
};

}  // namespace lk
