#pragma once
#include <ostream>
#include "Teuchos_ParameterList.hpp"
namespace Teuchos {
inline void writeParameterListToXmlOStream(const ParameterList& p, std::ostream& os, int indent = 0) {
  const std::string pad(indent, ' ');
  os << pad << "<ParameterList name=\"" << p.name() << "\">\n";
  for (auto& e : p.entries())
    os << pad << "  <Parameter name=\"" << e.name << "\" type=\"" << e.type << "\" value=\"" << e.value << "\"/>\n";
  for (auto& s : p.sublists()) writeParameterListToXmlOStream(*s, os, indent + 2);
  os << pad << "</ParameterList>\n";
}
}  // namespace Teuchos
