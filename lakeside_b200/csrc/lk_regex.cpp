#include "lk_regex.h"

#include <algorithm>
#include <cstring>

namespace lk {

struct Regex::Node {
  enum Kind { Empty, Char, Any, Class, Cat, Alt, Repeat, Bol, Eol, WordB, NWordB } kind;
  uint32_t cp = 0;       // Char: code point, Class: class index
  bool negated = false;  // Class
  int a = -1, b = -1;    // children
  int rmin = 0, rmax = -1;  // Repeat ({min,max}; max = -1 => unbounded)
};

static inline bool is_word(uint32_t c) {
  return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
}
static inline uint32_t fold_cp(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
static inline uint32_t swapcase(uint32_t c) {
  if (c >= 'A' && c <= 'Z') return c + 32;
  if (c >= 'a' && c <= 'z') return c - 32;
  return c;
}

static void decode_utf8(const char* s, size_t n, std::vector<uint32_t>& out) {
  out.clear();
  out.reserve(n);
  size_t i = 0;
  while (i < n) {
    unsigned char c = (unsigned char)s[i];
    uint32_t cp;
    int len;
    if (c < 0x80) { cp = c; len = 1; }
    else if ((c >> 5) == 6 && i + 1 < n) { cp = ((c & 0x1F) << 6) | (s[i + 1] & 0x3F); len = 2; }
    else if ((c >> 4) == 14 && i + 2 < n) { cp = ((c & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F); len = 3; }
    else if ((c >> 3) == 30 && i + 3 < n) { cp = ((c & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) | ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F); len = 4; }
    else { cp = 0xFFFD; len = 1; }
    out.push_back(cp);
    i += len;
  }
}

int Regex::new_node(int kind) {
  Node n;
  n.kind = (Node::Kind)kind;
  nodes_.push_back(n);
  return (int)nodes_.size() - 1;
}

uint32_t Regex::next_cp() {
  std::vector<uint32_t> tmp;
  unsigned char c = (unsigned char)cur();
  int len = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
  if (pos_ + len > pat_->size()) len = 1;
  decode_utf8(pat_->data() + pos_, len, tmp);
  pos_ += len;
  return tmp.empty() ? 0xFFFD : tmp[0];
}

int Regex::add_class(std::vector<Range> r, bool negated) {
  classes_.push_back(std::move(r));
  int n = new_node(Node::Class);
  nodes_[n].cp = (uint32_t)classes_.size() - 1;
  nodes_[n].negated = negated;
  return n;
}

static void posix_class(const std::string& name, std::vector<std::pair<uint32_t, uint32_t>>& out) {
  auto add = [&](uint32_t a, uint32_t b) { out.emplace_back(a, b); };
  if (name == "alpha") { add('a', 'z'); add('A', 'Z'); }
  else if (name == "digit") add('0', '9');
  else if (name == "alnum") { add('a', 'z'); add('A', 'Z'); add('0', '9'); }
  else if (name == "upper") add('A', 'Z');
  else if (name == "lower") add('a', 'z');
  else if (name == "space") { add('\t', '\r'); add(' ', ' '); }
  else if (name == "blank") { add(' ', ' '); add('\t', '\t'); }
  else if (name == "punct") { add('!', '/'); add(':', '@'); add('[', '`'); add('{', '~'); }
  else if (name == "xdigit") { add('0', '9'); add('a', 'f'); add('A', 'F'); }
  else if (name == "word") { add('a', 'z'); add('A', 'Z'); add('0', '9'); add('_', '_'); }
  else if (name == "cntrl") { add(0, 31); add(127, 127); }
  else if (name == "print") add(' ', '~');
  else if (name == "graph") add('!', '~');
  else fail(LK_ERR_UNSUPPORTED, "regex: unknown POSIX class [:" + name + ":]");
}

// \d \w \s (and negations) -> ranges.  Returns false if the escape is not a class escape.
bool Regex::parse_escape_class(std::vector<Range>& out, bool& negated) {
  char c = cur();
  negated = false;
  switch (c) {
    case 'D': negated = true; [[fallthrough]];
    case 'd': out.push_back({'0', '9'}); break;
    case 'W': negated = true; [[fallthrough]];
    case 'w': out.push_back({'0', '9'}); out.push_back({'A', 'Z'}); out.push_back({'_', '_'}); out.push_back({'a', 'z'}); break;
    case 'S': negated = true; [[fallthrough]];
    case 's': out.push_back({'\t', '\n'}); out.push_back({'\f', '\r'}); out.push_back({' ', ' '}); break;
    default: return false;
  }
  pos_++;
  return true;
}

uint32_t Regex::parse_escape_char() {
  LK_CHECK(!eof(), LK_ERR_UNSUPPORTED, "regex: trailing backslash");
  char c = cur();
  switch (c) {
    case 't': pos_++; return '\t';
    case 'n': pos_++; return '\n';
    case 'r': pos_++; return '\r';
    case 'f': pos_++; return '\f';
    case 'v': pos_++; return '\v';
    case 'a': pos_++; return 7;
    case 'x': {
      pos_++;
      uint32_t v = 0;
      auto hex = [&](char h) -> int {
        if (h >= '0' && h <= '9') return h - '0';
        if (h >= 'a' && h <= 'f') return h - 'a' + 10;
        if (h >= 'A' && h <= 'F') return h - 'A' + 10;
        return -1;
      };
      if (!eof() && cur() == '{') {
        pos_++;
        while (!eof() && cur() != '}') {
          int h = hex(cur());
          LK_CHECK(h >= 0, LK_ERR_UNSUPPORTED, "regex: bad \\x{...}");
          v = v * 16 + h;
          pos_++;
        }
        LK_CHECK(!eof(), LK_ERR_UNSUPPORTED, "regex: bad \\x{...}");
        pos_++;
      } else {
        for (int i = 0; i < 2; i++) {
          LK_CHECK(!eof() && hex(cur()) >= 0, LK_ERR_UNSUPPORTED, "regex: bad \\xHH");
          v = v * 16 + hex(cur());
          pos_++;
        }
      }
      return v;
    }
    default:
      if ((c >= '0' && c <= '9')) fail(LK_ERR_UNSUPPORTED, "regex: back-references are not supported (nor by RE2)");
      if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) fail(LK_ERR_UNSUPPORTED, std::string("regex: unsupported escape \\") + c);
      return next_cp();  // escaped punctuation
  }
}

int Regex::parse_class() {
  // cur() is just past '['
  std::vector<Range> r;
  bool negated = false;
  if (!eof() && cur() == '^') { negated = true; pos_++; }
  bool first = true;
  while (true) {
    LK_CHECK(!eof(), LK_ERR_UNSUPPORTED, "regex: missing ]");
    if (cur() == ']' && !first) { pos_++; break; }
    first = false;
    if (cur() == '[' && pos_ + 1 < pat_->size() && (*pat_)[pos_ + 1] == ':') {
      size_t e = pat_->find(":]", pos_ + 2);
      LK_CHECK(e != std::string::npos, LK_ERR_UNSUPPORTED, "regex: bad [:class:]");
      std::string name = pat_->substr(pos_ + 2, e - pos_ - 2);
      bool neg = false;
      if (!name.empty() && name[0] == '^') { neg = true; name = name.substr(1); }
      LK_CHECK(!neg, LK_ERR_UNSUPPORTED, "regex: negated POSIX class inside []");
      std::vector<std::pair<uint32_t, uint32_t>> pr;
      posix_class(name, pr);
      for (auto& p : pr) r.push_back({p.first, p.second});
      pos_ = e + 2;
      continue;
    }
    uint32_t lo;
    if (cur() == '\\') {
      pos_++;
      LK_CHECK(!eof(), LK_ERR_UNSUPPORTED, "regex: trailing backslash");
      std::vector<Range> er;
      bool eneg;
      if (parse_escape_class(er, eneg)) {
        LK_CHECK(!eneg, LK_ERR_UNSUPPORTED, "regex: negated class escape inside []");
        for (auto& x : er) r.push_back(x);
        continue;
      }
      lo = parse_escape_char();
    } else {
      lo = next_cp();
    }
    uint32_t hi = lo;
    if (pos_ + 1 < pat_->size() && cur() == '-' && (*pat_)[pos_ + 1] != ']') {
      pos_++;
      if (cur() == '\\') { pos_++; hi = parse_escape_char(); }
      else hi = next_cp();
      LK_CHECK(hi >= lo, LK_ERR_UNSUPPORTED, "regex: bad character range");
    }
    r.push_back({lo, hi});
  }
  return add_class(std::move(r), negated);
}

int Regex::parse_atom() {
  char c = cur();
  if (c == '(') {
    pos_++;
    if (!eof() && cur() == '?') {
      pos_++;
      LK_CHECK(!eof(), LK_ERR_UNSUPPORTED, "regex: bad group");
      if (cur() == ':') pos_++;
      else if (cur() == 'P' || cur() == '<') {  // (?P<name>...) / (?<name>...)
        if (cur() == 'P') pos_++;
        LK_CHECK(!eof() && cur() == '<', LK_ERR_UNSUPPORTED, "regex: bad named group");
        size_t e = pat_->find('>', pos_);
        LK_CHECK(e != std::string::npos, LK_ERR_UNSUPPORTED, "regex: bad named group");
        LK_CHECK(pos_ + 1 >= pat_->size() || ((*pat_)[pos_ + 1] != '=' && (*pat_)[pos_ + 1] != '!'), LK_ERR_UNSUPPORTED,
                 "regex: look-behind is not supported (nor by RE2)");
        pos_ = e + 1;
      } else if (cur() == '=' || cur() == '!') {
        fail(LK_ERR_UNSUPPORTED, "regex: look-ahead is not supported (nor by RE2)");
      } else {
        // flags: only i / s-less subset; applies to the whole pattern (documented deviation for mid-pattern flags)
        bool any = false;
        while (!eof() && cur() != ')' && cur() != ':') {
          LK_CHECK(cur() == 'i', LK_ERR_UNSUPPORTED, std::string("regex: unsupported flag ") + cur());
          fold_ = true;
          any = true;
          pos_++;
        }
        LK_CHECK(any && !eof(), LK_ERR_UNSUPPORTED, "regex: bad flag group");
        if (cur() == ')') { pos_++; return new_node(Node::Empty); }
        pos_++;  // ':'
      }
    }
    int n = parse_alt();
    LK_CHECK(!eof() && cur() == ')', LK_ERR_UNSUPPORTED, "regex: missing )");
    pos_++;
    return n;
  }
  if (c == '[') { pos_++; return parse_class(); }
  if (c == '.') { pos_++; return new_node(Node::Any); }
  if (c == '^') { pos_++; return new_node(Node::Bol); }
  if (c == '$') { pos_++; return new_node(Node::Eol); }
  if (c == '\\') {
    pos_++;
    LK_CHECK(!eof(), LK_ERR_UNSUPPORTED, "regex: trailing backslash");
    char e = cur();
    if (e == 'b') { pos_++; return new_node(Node::WordB); }
    if (e == 'B') { pos_++; return new_node(Node::NWordB); }
    if (e == 'A') { pos_++; return new_node(Node::Bol); }
    if (e == 'z') { pos_++; return new_node(Node::Eol); }
    if (e == 'Q') {  // \Q...\E literal
      pos_++;
      int acc = new_node(Node::Empty);
      while (!eof() && !(cur() == '\\' && pos_ + 1 < pat_->size() && (*pat_)[pos_ + 1] == 'E')) {
        int ch = new_node(Node::Char);
        nodes_[ch].cp = next_cp();
        int cat = new_node(Node::Cat);
        nodes_[cat].a = acc;
        nodes_[cat].b = ch;
        acc = cat;
      }
      if (!eof()) pos_ += 2;
      return acc;
    }
    std::vector<Range> er;
    bool neg;
    if (parse_escape_class(er, neg)) return add_class(std::move(er), neg);
    int n = new_node(Node::Char);
    nodes_[n].cp = parse_escape_char();
    return n;
  }
  LK_CHECK(c != '*' && c != '+' && c != '?', LK_ERR_UNSUPPORTED, "regex: missing argument to repetition operator");
  int n = new_node(Node::Char);
  nodes_[n].cp = next_cp();
  return n;
}

int Regex::parse_repeat() {
  int atom = parse_atom();
  while (!eof()) {
    char c = cur();
    int rmin, rmax;
    if (c == '*') { rmin = 0; rmax = -1; pos_++; }
    else if (c == '+') { rmin = 1; rmax = -1; pos_++; }
    else if (c == '?') { rmin = 0; rmax = 1; pos_++; }
    else if (c == '{') {
      // {n} {n,} {n,m}; anything else is a literal '{' (RE2 behaviour)
      size_t q = pos_ + 1;
      auto num = [&](int& v) {
        if (q >= pat_->size() || !isdigit((unsigned char)(*pat_)[q])) return false;
        v = 0;
        while (q < pat_->size() && isdigit((unsigned char)(*pat_)[q])) { v = v * 10 + ((*pat_)[q] - '0'); q++; if (v > 1000) return false; }
        return true;
      };
      int a, b = -2;
      if (!num(a)) break;
      if (q < pat_->size() && (*pat_)[q] == ',') {
        q++;
        if (q < pat_->size() && (*pat_)[q] == '}') b = -1;
        else if (!num(b)) break;
      } else b = a;
      if (q >= pat_->size() || (*pat_)[q] != '}') break;
      LK_CHECK(b == -1 || b >= a, LK_ERR_UNSUPPORTED, "regex: bad repetition {n,m}");
      rmin = a; rmax = b; pos_ = q + 1;
    } else break;
    if (!eof() && (cur() == '?')) pos_++;  // lazy: irrelevant for a boolean match
    LK_CHECK(eof() || cur() != '+', LK_ERR_UNSUPPORTED, "regex: possessive quantifiers are not supported");
    int r = new_node(Node::Repeat);
    nodes_[r].a = atom;
    nodes_[r].rmin = rmin;
    nodes_[r].rmax = rmax;
    atom = r;
  }
  return atom;
}

int Regex::parse_concat() {
  int acc = -1;
  while (!eof() && cur() != '|' && cur() != ')') {
    int n = parse_repeat();
    if (acc < 0) acc = n;
    else {
      int cat = new_node(Node::Cat);
      nodes_[cat].a = acc;
      nodes_[cat].b = n;
      acc = cat;
    }
  }
  return acc < 0 ? new_node(Node::Empty) : acc;
}

int Regex::parse_alt() {
  int acc = parse_concat();
  while (!eof() && cur() == '|') {
    pos_++;
    int rhs = parse_concat();
    int alt = new_node(Node::Alt);
    nodes_[alt].a = acc;
    nodes_[alt].b = rhs;
    acc = alt;
  }
  return acc;
}

void Regex::emit(int ni) {
  LK_CHECK(prog_.size() < (1u << 20), LK_ERR_UNSUPPORTED, "regex: program too large");
  const Node n = nodes_[ni];
  switch (n.kind) {
    case Node::Empty: break;
    case Node::Char: { Inst i; i.op = Inst::Char; i.x = n.cp; prog_.push_back(i); break; }
    case Node::Any: { Inst i; i.op = Inst::Any; prog_.push_back(i); break; }
    case Node::Class: { Inst i; i.op = Inst::Class; i.x = n.cp; i.y = n.negated; prog_.push_back(i); break; }
    case Node::Bol: { Inst i; i.op = Inst::Bol; prog_.push_back(i); break; }
    case Node::Eol: { Inst i; i.op = Inst::Eol; prog_.push_back(i); break; }
    case Node::WordB: { Inst i; i.op = Inst::WordB; prog_.push_back(i); break; }
    case Node::NWordB: { Inst i; i.op = Inst::NWordB; prog_.push_back(i); break; }
    case Node::Cat: emit(n.a); emit(n.b); break;
    case Node::Alt: {
      size_t split = prog_.size();
      prog_.push_back(Inst{Inst::Split, 0, 0});
      prog_[split].x = (uint32_t)prog_.size();
      emit(n.a);
      size_t jmp = prog_.size();
      prog_.push_back(Inst{Inst::Jmp, 0, 0});
      prog_[split].y = (uint32_t)prog_.size();
      emit(n.b);
      prog_[jmp].x = (uint32_t)prog_.size();
      break;
    }
    case Node::Repeat: {
      for (int i = 0; i < n.rmin; i++) emit(n.a);
      if (n.rmax < 0) {
        size_t l0 = prog_.size();
        prog_.push_back(Inst{Inst::Split, 0, 0});
        prog_[l0].x = (uint32_t)prog_.size();
        emit(n.a);
        prog_.push_back(Inst{Inst::Jmp, (uint32_t)l0, 0});
        prog_[l0].y = (uint32_t)prog_.size();
      } else {
        std::vector<size_t> splits;
        for (int i = n.rmin; i < n.rmax; i++) {
          splits.push_back(prog_.size());
          prog_.push_back(Inst{Inst::Split, 0, 0});
          prog_[splits.back()].x = (uint32_t)prog_.size();
          emit(n.a);
        }
        for (size_t s : splits) prog_[s].y = (uint32_t)prog_.size();
      }
      break;
    }
  }
}

Regex::Regex(const std::string& pattern, bool case_insensitive) : fold_(case_insensitive) {
  pat_ = &pattern;
  pos_ = 0;
  int root = parse_alt();
  LK_CHECK(eof(), LK_ERR_UNSUPPORTED, "regex: unmatched )");
  emit(root);
  prog_.push_back(Inst{Inst::Match, 0, 0});
  if (fold_)
    for (auto& i : prog_)
      if (i.op == Inst::Char) i.x = fold_cp(i.x);
  pat_ = nullptr;
  nodes_.clear();
}

Regex::~Regex() = default;

bool Regex::class_match(const Inst& in, uint32_t cp) const {
  const auto& r = classes_[in.x];
  auto hit = [&](uint32_t c) {
    for (auto& x : r)
      if (c >= x.lo && c <= x.hi) return true;
    return false;
  };
  bool m = hit(cp) || (fold_ && swapcase(cp) != cp && hit(swapcase(cp)));
  return in.y ? !m : m;
}

void Regex::add_thread(std::vector<uint32_t>& list, std::vector<uint32_t>& mark, uint32_t gen, uint32_t pc, bool at_start,
                       bool at_end, bool prev_word, bool next_word) const {
  // iterative DFS over epsilon transitions
  uint32_t stack[64];
  std::vector<uint32_t> big;
  int sp = 0;
  auto push = [&](uint32_t v) {
    if (sp < 64) stack[sp++] = v;
    else big.push_back(v);
  };
  push(pc);
  while (sp > 0 || !big.empty()) {
    uint32_t p;
    if (!big.empty()) { p = big.back(); big.pop_back(); }
    else p = stack[--sp];
    if (mark[p] == gen) continue;
    mark[p] = gen;
    const Inst& in = prog_[p];
    switch (in.op) {
      case Inst::Jmp: push(in.x); break;
      case Inst::Split: push(in.y); push(in.x); break;
      case Inst::Bol: if (at_start) push(p + 1); break;
      case Inst::Eol: if (at_end) push(p + 1); break;
      case Inst::WordB: if (prev_word != next_word) push(p + 1); break;
      case Inst::NWordB: if (prev_word == next_word) push(p + 1); break;
      default: list.push_back(p); break;
    }
  }
}

bool Regex::search(const char* s, size_t n) const {
  std::vector<uint32_t> cps;
  decode_utf8(s, n, cps);
  const size_t L = cps.size();
  std::vector<uint32_t> clist, nlist, mark(prog_.size(), 0);
  uint32_t gen = 0;
  for (size_t i = 0; i <= L; i++) {
    bool at_start = i == 0, at_end = i == L;
    bool pw = i > 0 && is_word(cps[i - 1]);
    bool nw = i < L && is_word(cps[i]);
    if (i == 0) gen++;
    // unanchored search: a fresh thread starts at every position (marks of this generation are shared with the
    // threads carried over from the previous position, which were added with the same context)
    add_thread(clist, mark, gen, 0, at_start, at_end, pw, nw);
    if (clist.empty() && at_end) return false;
    uint32_t cp = i < L ? cps[i] : 0;
    uint32_t fcp = fold_ ? fold_cp(cp) : cp;
    gen++;
    bool npw = i < L && is_word(cp);
    bool nnw = i + 1 < L && is_word(cps[i + 1]);
    for (uint32_t p : clist) {
      const Inst& in = prog_[p];
      if (in.op == Inst::Match) return true;
      if (at_end) continue;
      bool ok = false;
      if (in.op == Inst::Char) ok = in.x == fcp;
      else if (in.op == Inst::Any) ok = cp != '\n';
      else if (in.op == Inst::Class) ok = class_match(in, cp);
      if (ok) add_thread(nlist, mark, gen, p + 1, false, i + 1 == L, npw, nnw);
    }
    clist.swap(nlist);
    nlist.clear();
  }
  return false;
}

}  // namespace lk
