// xml_mini.h — minimal XML DOM reader (elements, attributes, comments, declarations).
// Enough for MJCF; no entities beyond &lt; &gt; &amp; &quot; &apos;, no CDATA, no namespaces.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace b2mj {

struct XmlNode {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attrs;  // in document order
  std::vector<std::unique_ptr<XmlNode>> children;
  int line = 0;

  const std::string* attr(const std::string& k) const {
    for (auto& a : attrs)
      if (a.first == k) return &a.second;
    return nullptr;
  }
  const XmlNode* child(const std::string& t) const {
    for (auto& c : children)
      if (c->tag == t) return c.get();
    return nullptr;
  }
};

// Parses `text`; on failure returns nullptr and fills err ("line N: message").
std::unique_ptr<XmlNode> xml_parse(const std::string& text, std::string& err);

}  // namespace b2mj
