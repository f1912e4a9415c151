#include "params.hpp"

#include <cstdio>

#include <cstdlib>
#include <cstring>

#include "../../include/hymls_b200.h"

namespace hymls {

static Error typeError(const std::string& n) {
  return Error(HYMLS_B200_ERR_ARG, "parameter '" + n + "' has the wrong type");
}

int ParameterList::get(const std::string& n, int def) {
  if (!isParameter(n)) set(n, def);
  const Value& v = vals_[n];
  if (v.kind == INT) return (int)v.i;
  if (v.kind == BOOL) return v.b ? 1 : 0;
  throw typeError(n);
}
bool ParameterList::get(const std::string& n, bool def) {
  if (!isParameter(n)) set(n, def);
  const Value& v = vals_[n];
  if (v.kind == BOOL) return v.b;
  if (v.kind == INT) return v.i != 0;
  throw typeError(n);
}
double ParameterList::get(const std::string& n, double def) {
  if (!isParameter(n)) set(n, def);
  const Value& v = vals_[n];
  if (v.kind == DOUBLE) return v.d;
  if (v.kind == INT) return (double)v.i;
  throw typeError(n);
}
std::string ParameterList::get(const std::string& n, const char* def) {
  if (!isParameter(n)) set(n, def);
  const Value& v = vals_[n];
  if (v.kind == STRING) return v.s;
  throw typeError(n);
}
void ParameterList::set(const std::string& n, int v) {
  Value& s = slot(n);
  s = Value();
  s.kind = INT;
  s.i = v;
}
void ParameterList::set(const std::string& n, bool v) {
  Value& s = slot(n);
  s = Value();
  s.kind = BOOL;
  s.b = v;
}
void ParameterList::set(const std::string& n, double v) {
  Value& s = slot(n);
  s = Value();
  s.kind = DOUBLE;
  s.d = v;
}
void ParameterList::set(const std::string& n, const char* v) {
  Value& s = slot(n);
  s = Value();
  s.kind = STRING;
  s.s = v;
}

ParameterList ParameterList::deepCopy() const {
  ParameterList out;
  out.vals_ = vals_;
  out.order_ = order_;
  for (auto& kv : subs_) out.subs_[kv.first] = std::make_shared<ParameterList>(kv.second->deepCopy());
  return out;
}
std::vector<std::string> ParameterList::parameterNames() const {
  std::vector<std::string> r;
  for (auto& kv : vals_) r.push_back(kv.first);
  return r;
}
std::vector<std::string> ParameterList::sublistNames() const {
  std::vector<std::string> r;
  for (auto& kv : subs_) r.push_back(kv.first);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Minimal XML reader: elements ParameterList / Parameter with attributes; comments skipped.
// ---------------------------------------------------------------------------------------------
namespace {
struct Parser {
  const std::string& s;
  size_t p = 0;
  explicit Parser(const std::string& str) : s(str) {}

  void skipWsAndComments() {
    for (;;) {
      while (p < s.size() && isspace((unsigned char)s[p])) ++p;
      if (s.compare(p, 4, "<!--") == 0) {
        size_t e = s.find("-->", p + 4);
        if (e == std::string::npos) throw Error(HYMLS_B200_ERR_ARG, "XML: unterminated comment");
        p = e + 3;
      } else if (s.compare(p, 2, "<?") == 0) {
        size_t e = s.find("?>", p + 2);
        if (e == std::string::npos) throw Error(HYMLS_B200_ERR_ARG, "XML: unterminated declaration");
        p = e + 2;
      } else {
        return;
      }
    }
  }
  static std::string unescape(const std::string& v) {
    std::string o;
    for (size_t i = 0; i < v.size(); ++i) {
      if (v[i] == '&') {
        if (v.compare(i, 4, "&lt;") == 0) { o += '<'; i += 3; continue; }
        if (v.compare(i, 4, "&gt;") == 0) { o += '>'; i += 3; continue; }
        if (v.compare(i, 5, "&amp;") == 0) { o += '&'; i += 4; continue; }
        if (v.compare(i, 6, "&quot;") == 0) { o += '"'; i += 5; continue; }
        if (v.compare(i, 6, "&apos;") == 0) { o += '\''; i += 5; continue; }
      }
      o += v[i];
    }
    return o;
  }
  // parses "<tag a="b" ...>" or "<tag .../>"; returns tag, fills attrs, selfClosing
  std::string openTag(std::map<std::string, std::string>& attrs, bool& selfClosing) {
    if (s[p] != '<') throw Error(HYMLS_B200_ERR_ARG, "XML: expected '<'");
    ++p;
    size_t b = p;
    while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/') ++p;
    std::string tag = s.substr(b, p - b);
    for (;;) {
      while (p < s.size() && isspace((unsigned char)s[p])) ++p;
      if (p >= s.size()) throw Error(HYMLS_B200_ERR_ARG, "XML: unterminated tag");
      if (s[p] == '/') {
        selfClosing = true;
        p += 2;
        return tag;
      }
      if (s[p] == '>') {
        selfClosing = false;
        ++p;
        return tag;
      }
      size_t nb = p;
      while (p < s.size() && s[p] != '=' && !isspace((unsigned char)s[p])) ++p;
      std::string name = s.substr(nb, p - nb);
      while (p < s.size() && (isspace((unsigned char)s[p]) || s[p] == '=')) ++p;
      char q = s[p];
      if (q != '"' && q != '\'') throw Error(HYMLS_B200_ERR_ARG, "XML: attribute value must be quoted");
      ++p;
      size_t vb = p;
      while (p < s.size() && s[p] != q) ++p;
      attrs[name] = unescape(s.substr(vb, p - vb));
      ++p;
    }
  }
  void parseList(ParameterList& pl) {
    for (;;) {
      skipWsAndComments();
      if (p >= s.size()) throw Error(HYMLS_B200_ERR_ARG, "XML: missing </ParameterList>");
      if (s.compare(p, 2, "</") == 0) {
        size_t e = s.find('>', p);
        p = e + 1;
        return;
      }
      std::map<std::string, std::string> a;
      bool sc = false;
      std::string tag = openTag(a, sc);
      if (tag == "ParameterList") {
        ParameterList& sub = pl.sublist(a["name"]);
        if (!sc) parseList(sub);
      } else if (tag == "Parameter") {
        const std::string &n = a["name"], &t = a["type"], &v = a["value"];
        if (t == "bool") {
          pl.set(n, v == "1" || v == "true" || v == "True" || v == "TRUE");
        } else if (t == "int" || t == "long long" || t == "long") {
          pl.set(n, (int)strtoll(v.c_str(), nullptr, 10));
        } else if (t == "double" || t == "float") {
          pl.set(n, strtod(v.c_str(), nullptr));
        } else {
          pl.set(n, v.c_str());
        }
        if (!sc) {  // <Parameter ...></Parameter>
          skipWsAndComments();
          size_t e = s.find('>', p);
          p = e + 1;
        }
      } else {
        throw Error(HYMLS_B200_ERR_ARG, "XML: unexpected element <" + tag + ">");
      }
    }
  }
};
}  // namespace

ParameterList ParameterList::fromXml(const std::string& xml) {
  Parser ps(xml);
  ps.skipWsAndComments();
  std::map<std::string, std::string> a;
  bool sc = false;
  std::string tag = ps.openTag(a, sc);
  if (tag != "ParameterList") throw Error(HYMLS_B200_ERR_ARG, "XML: root element must be ParameterList");
  ParameterList root;
  if (!sc) ps.parseList(root);
  return root;
}

static std::string xmlEscape(const std::string& t) {
  std::string o;
  for (char c : t) {
    switch (c) {
      case '&': o += "&amp;"; break;
      case '<': o += "&lt;"; break;
      case '>': o += "&gt;"; break;
      case '"': o += "&quot;"; break;
      default: o += c;
    }
  }
  return o;
}

std::string ParameterList::toXml(const std::string& name, int indent) const {
  const std::string pad((size_t)indent, ' ');
  std::string o = pad + "<ParameterList name=\"" + xmlEscape(name) + "\">\n";
  for (const std::string& n : order_) {
    auto v = vals_.find(n);
    if (v != vals_.end()) {
      const Value& x = v->second;
      char buf[64];
      std::string type, val;
      switch (x.kind) {
        case BOOL: type = "bool"; val = x.b ? "true" : "false"; break;
        case INT: type = "int"; val = std::to_string(x.i); break;
        case DOUBLE: type = "double"; snprintf(buf, sizeof buf, "%.17g", x.d); val = buf; break;
        default: type = "string"; val = xmlEscape(x.s);
      }
      o += pad + "  <Parameter name=\"" + xmlEscape(n) + "\" type=\"" + type + "\" value=\"" + val + "\"/>\n";
      continue;
    }
    auto sub = subs_.find(n);
    if (sub != subs_.end()) o += sub->second->toXml(n, indent + 2);
  }
  o += pad + "</ParameterList>\n";
  return o;
}

}  // namespace hymls
