#include "xml_mini.h"

#include <cctype>

namespace b2mj {
namespace {

struct Cursor {
  const std::string& s;
  size_t i = 0;
  int line = 1;
  explicit Cursor(const std::string& t) : s(t) {}
  bool eof() const { return i >= s.size(); }
  char peek() const { return s[i]; }
  void adv() {
    if (s[i] == '\n') line++;
    i++;
  }
  bool starts(const char* lit) const { return s.compare(i, std::char_traits<char>::length(lit), lit) == 0; }
  void skip_ws() {
    while (!eof() && std::isspace((unsigned char)peek())) adv();
  }
  bool skip_until(const char* lit) {
    while (!eof()) {
      if (starts(lit)) {
        for (size_t k = 0; lit[k]; k++) adv();
        return true;
      }
      adv();
    }
    return false;
  }
};

std::string unescape(const std::string& v) {
  std::string o;
  o.reserve(v.size());
  for (size_t i = 0; i < v.size(); i++) {
    if (v[i] == '&') {
      static const std::pair<const char*, char> ents[] = {
          {"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
      bool hit = false;
      for (auto& e : ents) {
        size_t n = std::char_traits<char>::length(e.first);
        if (v.compare(i, n, e.first) == 0) {
          o.push_back(e.second);
          i += n - 1;
          hit = true;
          break;
        }
      }
      if (hit) continue;
    }
    o.push_back(v[i]);
  }
  return o;
}

bool is_name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }

// skip comments, processing instructions, doctype and text between elements
bool skip_misc(Cursor& c, std::string& err) {
  for (;;) {
    c.skip_ws();
    if (c.eof()) return true;
    if (c.starts("<!--")) {
      if (!c.skip_until("-->")) {
        err = "line " + std::to_string(c.line) + ": unterminated comment";
        return false;
      }
    } else if (c.starts("<?")) {
      if (!c.skip_until("?>")) {
        err = "line " + std::to_string(c.line) + ": unterminated declaration";
        return false;
      }
    } else if (c.starts("<!")) {
      if (!c.skip_until(">")) {
        err = "line " + std::to_string(c.line) + ": unterminated <! block";
        return false;
      }
    } else if (c.peek() != '<') {
      c.adv();  // free text is ignored (MJCF carries no text content)
    } else {
      return true;
    }
  }
}

std::unique_ptr<XmlNode> parse_element(Cursor& c, std::string& err) {
  // precondition: c.peek()=='<' and next is a name char
  int line = c.line;
  c.adv();
  std::string tag;
  while (!c.eof() && is_name_char(c.peek())) {
    tag.push_back(c.peek());
    c.adv();
  }
  if (tag.empty()) {
    err = "line " + std::to_string(line) + ": expected element name";
    return nullptr;
  }
  auto node = std::make_unique<XmlNode>();
  node->tag = tag;
  node->line = line;
  for (;;) {
    c.skip_ws();
    if (c.eof()) {
      err = "line " + std::to_string(line) + ": unterminated element <" + tag + ">";
      return nullptr;
    }
    if (c.starts("/>")) {
      c.adv();
      c.adv();
      return node;
    }
    if (c.peek() == '>') {
      c.adv();
      break;
    }
    std::string key;
    while (!c.eof() && is_name_char(c.peek())) {
      key.push_back(c.peek());
      c.adv();
    }
    c.skip_ws();
    if (key.empty() || c.eof() || c.peek() != '=') {
      err = "line " + std::to_string(c.line) + ": malformed attribute in <" + tag + ">";
      return nullptr;
    }
    c.adv();
    c.skip_ws();
    if (c.eof() || (c.peek() != '"' && c.peek() != '\'')) {
      err = "line " + std::to_string(c.line) + ": attribute value must be quoted";
      return nullptr;
    }
    char q = c.peek();
    c.adv();
    std::string val;
    while (!c.eof() && c.peek() != q) {
      val.push_back(c.peek());
      c.adv();
    }
    if (c.eof()) {
      err = "line " + std::to_string(c.line) + ": unterminated attribute value";
      return nullptr;
    }
    c.adv();
    node->attrs.emplace_back(key, unescape(val));
  }
  // children until </tag>
  for (;;) {
    if (!skip_misc(c, err)) return nullptr;
    if (c.eof()) {
      err = "line " + std::to_string(line) + ": missing </" + tag + ">";
      return nullptr;
    }
    if (c.starts("</")) {
      c.adv();
      c.adv();
      std::string close;
      while (!c.eof() && is_name_char(c.peek())) {
        close.push_back(c.peek());
        c.adv();
      }
      c.skip_ws();
      if (c.eof() || c.peek() != '>' || close != tag) {
        err = "line " + std::to_string(c.line) + ": mismatched </" + close + "> for <" + tag + ">";
        return nullptr;
      }
      c.adv();
      return node;
    }
    auto ch = parse_element(c, err);
    if (!ch) return nullptr;
    node->children.push_back(std::move(ch));
  }
}

}  // namespace

std::unique_ptr<XmlNode> xml_parse(const std::string& text, std::string& err) {
  Cursor c(text);
  if (!skip_misc(c, err)) return nullptr;
  if (c.eof()) {
    err = "empty document";
    return nullptr;
  }
  auto root = parse_element(c, err);
  if (!root) return nullptr;
  if (!skip_misc(c, err)) return nullptr;
  if (!c.eof()) {
    err = "line " + std::to_string(c.line) + ": content after root element";
    return nullptr;
  }
  return root;
}

}  // namespace b2mj
