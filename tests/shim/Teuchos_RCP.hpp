#pragma once
#include <memory>
#include <stdexcept>
namespace Teuchos {
struct ENull {};
static const ENull null = ENull();
template <class T>
class RCP {
 public:
  RCP() {}
  RCP(ENull) {}
  explicit RCP(T* p) : p_(p) {}
  RCP(const std::shared_ptr<T>& p) : p_(p) {}
  template <class U>
  RCP(const RCP<U>& o) : p_(o.shared()) {}
  T* operator->() const { return p_.get(); }
  T& operator*() const { return *p_; }
  T* get() const { return p_.get(); }
  bool operator==(ENull) const { return !p_; }
  bool operator!=(ENull) const { return (bool)p_; }
  const std::shared_ptr<T>& shared() const { return p_; }
 private:
  std::shared_ptr<T> p_;
};
template <class T>
RCP<T> rcp(T* p) { return RCP<T>(p); }
template <class T, class U>
RCP<T> rcp_dynamic_cast(const RCP<U>& p, bool throwOnFail = false) {
  std::shared_ptr<T> q = std::dynamic_pointer_cast<T>(p.shared());
  if (!q && throwOnFail) throw std::runtime_error("rcp_dynamic_cast failed");
  return RCP<T>(q);
}
}  // namespace Teuchos
