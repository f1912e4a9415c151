"""Teuchos::ParameterList stand-in (ORACLE -- test infrastructure only).

A ParameterList is a nested dict; `get(name, default)` writes the default back
like Teuchos does (reference relies on that, e.g. src/HYMLS_BasePartitioner.cpp:139-141).
XML subset: <ParameterList name=..> / <Parameter name type value/> with types
bool / int / double / string (testSuite/*.xml).
"""
import copy
import xml.etree.ElementTree as ET


class ParameterList(dict):
    def get(self, name, default=None):  # Teuchos semantics: default is stored
        if name not in self:
            if default is None:
                raise KeyError(name)
            self[name] = default
        return self[name]

    def set(self, name, value):
        self[name] = value

    def sublist(self, name):
        if name not in self:
            self[name] = ParameterList()
        return self[name]

    def isParameter(self, name):
        return name in self and not isinstance(self[name], ParameterList)

    def isSublist(self, name):
        return name in self and isinstance(self[name], ParameterList)

    def copy(self):
        return copy.deepcopy(self)


def _convert(typ, val):
    if typ == "bool":
        return val.strip().lower() in ("1", "true")
    if typ == "int":
        return int(val)
    if typ == "double":
        return float(val)
    return val


def _from_elem(elem):
    pl = ParameterList()
    for ch in elem:
        if ch.tag == "ParameterList":
            pl[ch.attrib["name"]] = _from_elem(ch)
        elif ch.tag == "Parameter":
            pl[ch.attrib["name"]] = _convert(ch.attrib.get("type", "string"), ch.attrib["value"])
    return pl


def read_xml(path_or_string):
    if "<ParameterList" in path_or_string:
        root = ET.fromstring(path_or_string)
    else:
        root = ET.parse(path_or_string).getroot()
    return _from_elem(root)


def overlay(base, over):
    """Teuchos setParameters: entries of `over` replace / extend `base` recursively."""
    for k, v in over.items():
        if isinstance(v, ParameterList) and isinstance(base.get(k, None) if k in base else None, ParameterList):
            overlay(base[k], v)
        else:
            base[k] = copy.deepcopy(v)
    return base
