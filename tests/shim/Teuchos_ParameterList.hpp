#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
namespace Teuchos {
// nested name -> (type, value) list with insertion order, the subset of Teuchos::ParameterList the adapter and the
// driver program use
class ParameterList {
 public:
  explicit ParameterList(const std::string& name = "ANONYMOUS") : name_(name) {}
  const std::string& name() const { return name_; }
  ParameterList& set(const std::string& n, int v) { return put(n, "int", std::to_string(v)); }
  ParameterList& set(const std::string& n, bool v) { return put(n, "bool", v ? "true" : "false"); }
  ParameterList& set(const std::string& n, double v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.17g", v);
    return put(n, "double", buf);
  }
  ParameterList& set(const std::string& n, const char* v) { return put(n, "string", v); }
  ParameterList& set(const std::string& n, const std::string& v) { return put(n, "string", v); }
  ParameterList& sublist(const std::string& n) {
    for (auto& s : subs_)
      if (s->name() == n) return *s;
    subs_.emplace_back(new ParameterList(n));
    return *subs_.back();
  }
  // entries of `other` overwrite / extend this list (sublists recursively)
  ParameterList& setParameters(const ParameterList& other) {
    for (auto& e : other.entries_) put(e.name, e.type, e.value);
    for (auto& s : other.subs_) sublist(s->name()).setParameters(*s);
    return *this;
  }
  struct Entry { std::string name, type, value; };
  const std::vector<Entry>& entries() const { return entries_; }
  const std::vector<std::shared_ptr<ParameterList>>& sublists() const { return subs_; }
 private:
  ParameterList& put(const std::string& n, const std::string& t, const std::string& v) {
    for (auto& e : entries_)
      if (e.name == n) { e.type = t; e.value = v; return *this; }
    entries_.push_back({n, t, v});
    return *this;
  }
  std::string name_;
  std::vector<Entry> entries_;
  std::vector<std::shared_ptr<ParameterList>> subs_;
};
}  // namespace Teuchos
