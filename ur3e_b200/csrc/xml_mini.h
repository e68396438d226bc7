// Minimal XML DOM reader for MJCF files (elements, attributes, comments, <?...?>).
// No entities/CDATA/namespaces: MJCF files in this repo do not use them.
#pragma once
#include <cctype>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ur3e {

struct XmlNode {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attrs;  // document order
  std::vector<std::unique_ptr<XmlNode>> children;

  const std::string* find(const std::string& k) const {
    for (auto& a : attrs) if (a.first == k) return &a.second;
    return nullptr;
  }
  const XmlNode* child(const std::string& t) const {
    for (auto& c : children) if (c->tag == t) return c.get();
    return nullptr;
  }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& text) : s_(text), p_(0) {}
  std::unique_ptr<XmlNode> parse() {
    skip_misc();
    auto root = element();
    if (!root) throw std::runtime_error("xml: no root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t p_;
  bool starts(const char* lit) const { return s_.compare(p_, std::char_traits<char>::length(lit), lit) == 0; }
  void skip_ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) ++p_; }
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (starts("<!--")) { size_t e = s_.find("-->", p_); if (e == std::string::npos) throw std::runtime_error("xml: open comment"); p_ = e + 3; }
      else if (starts("<?")) { size_t e = s_.find("?>", p_); if (e == std::string::npos) throw std::runtime_error("xml: open PI"); p_ = e + 2; }
      else if (starts("<!")) { size_t e = s_.find('>', p_); if (e == std::string::npos) throw std::runtime_error("xml: open decl"); p_ = e + 1; }
      else return;
    }
  }
  std::string name() {
    size_t b = p_;
    while (p_ < s_.size() && (std::isalnum((unsigned char)s_[p_]) || s_[p_] == '_' || s_[p_] == '-' || s_[p_] == ':' || s_[p_] == '.')) ++p_;
    if (b == p_) throw std::runtime_error("xml: expected name at offset " + std::to_string(p_));
    return s_.substr(b, p_ - b);
  }
  std::unique_ptr<XmlNode> element() {
    if (p_ >= s_.size() || s_[p_] != '<') return nullptr;
    ++p_;
    auto n = std::make_unique<XmlNode>();
    n->tag = name();
    for (;;) {
      skip_ws();
      if (p_ >= s_.size()) throw std::runtime_error("xml: unexpected end in <" + n->tag);
      if (s_[p_] == '/') { if (s_.compare(p_, 2, "/>") != 0) throw std::runtime_error("xml: bad '/'"); p_ += 2; return n; }
      if (s_[p_] == '>') { ++p_; break; }
      std::string k = name();
      skip_ws();
      if (s_[p_] != '=') throw std::runtime_error("xml: expected '=' after " + k);
      ++p_; skip_ws();
      char q = s_[p_];
      if (q != '"' && q != '\'') throw std::runtime_error("xml: expected quote");
      size_t e = s_.find(q, p_ + 1);
      if (e == std::string::npos) throw std::runtime_error("xml: open attribute");
      n->attrs.emplace_back(k, s_.substr(p_ + 1, e - p_ - 1));
      p_ = e + 1;
    }
    for (;;) {
      // text content is ignored
      while (p_ < s_.size() && s_[p_] != '<') ++p_;
      if (p_ >= s_.size()) throw std::runtime_error("xml: missing </" + n->tag + ">");
      if (starts("<!--") || starts("<?") || starts("<!")) { skip_misc(); continue; }
      if (starts("</")) {
        p_ += 2; std::string t = name(); skip_ws();
        if (t != n->tag || s_[p_] != '>') throw std::runtime_error("xml: mismatched </" + t + "> for <" + n->tag + ">");
        ++p_; return n;
      }
      n->children.push_back(element());
    }
  }
};

}  // namespace ur3e
