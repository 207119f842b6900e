"""Mirror C structs / enums of a header in ctypes by parsing it (keeps Python and C in step)."""
import ctypes as C
import re

_CT = {
    "int8_t": C.c_int8,
    "int16_t": C.c_int16,
    "int32_t": C.c_int32,
    "int64_t": C.c_int64,
    "double": C.c_double,
}


def _strip_comments(txt):
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def parse_struct(header_path, name):
    """ctypes `_fields_` of `typedef struct name { ... } name;` (one member per ';')."""
    txt = _strip_comments(open(header_path).read())
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), txt, re.S)
    if not m:
        raise KeyError(name)
    fields = []
    for line in m.group(1).split(";"):
        line = " ".join(line.split())
        if not line:
            continue
        mm = re.match(r"(const )?(\w+) ?(\*?) ?(\w+)(\[(\d+)\])?$", line)
        if not mm:
            raise ValueError("cannot parse struct member: %r" % line)
        base = _CT[mm.group(2)]
        if mm.group(3):
            ct = C.POINTER(base)
        elif mm.group(6):
            ct = base * int(mm.group(6))
        else:
            ct = base
        fields.append((mm.group(4), ct))
    return fields


def parse_enum(header_path, name):
    """dict name -> value of `enum name { A = 0, B, ... };`"""
    txt = _strip_comments(open(header_path).read())
    m = re.search(r"enum\s+%s\s*\{(.*?)\}\s*;" % name, txt, re.S)
    if not m:
        raise KeyError(name)
    out, val = {}, -1
    for item in m.group(1).split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = item.split("=")
            val = int(v.strip(), 0)
            out[k.strip()] = val
        else:
            val += 1
            out[item] = val
    return out


def declared_functions(header_path):
    """names of all `int|const char * name(` prototypes in the header"""
    txt = _strip_comments(open(header_path).read())
    return re.findall(r"^\s*(?:int|const char \*)\s*\*?(\w+)\s*\(", txt, re.M)
