"""chf2c.py -- TEST INFRASTRUCTURE.  A minimal, mechanical Chombo-Fortran (.ChF) -> C++ translator for exactly the
constructs the reference's two kernel files use (Source/VariableCoeffPoissonOperatorF.ChF, Source/SetLevelDataF.ChF).

Why: the image has no Fortran compiler, so the reference's kernels cannot be compiled.  Instead of trusting only the
hand-written restatement in mgic_oracle.cpp, `make -C oracle ref` runs this script over the reference's .ChF files WHERE
THEY LIE and compiles the output (oracle/_ref/gen/*.cpp, never committed) into oracle/_ref/libmgic_ref.so, under the
Fortran symbol names the reference's generated prototypes declare (gsrbhelmholtzvc3d_, ...).  Every arithmetic expression
is copied through token by token: the translator rewrites syntax only -- continuation lines, do/if blocks, the CHF_*
macros (3-D expansions of Chombo's ChF preprocessor: CHF_DTERM, CHF_IX, CHF_AUTOIX, CHF_OFFSETIX, CHF_MULTIDO, ...),
array declarations into index lambdas -- and it stops with an error on anything it does not know.  Expressions are
evaluated left to right as written (g++ -ffp-contract=off), which is what a Fortran compiler does without value-unsafe
optimisations; parentheses are kept.

usage: python chf2c.py file.ChF > file.cpp        (CH_SPACEDIM = 3)
"""
import re
import sys

SPACEDIM = 3
CONSTANTS = "static const double zero = 0.0, one = 1.0, two = 2.0, three = 3.0, four = 4.0, half = 0.5;"   # Chombo CONSTANTS.H


class ChfError(Exception):
    pass


def preprocess(text):
    """#if / #elif / #else / #endif on CH_SPACEDIM, #include dropped; comment lines dropped; continuation lines joined"""
    out, stack = [], []          # stack of [taken_before, active]
    for raw in text.splitlines():
        line = raw.rstrip()
        m = re.match(r"#\s*(if|elif|else|endif|include)\b(.*)", line)
        if m:
            kind, rest = m.group(1), m.group(2).strip()
            if kind == "include":
                continue
            if kind in ("if", "elif"):
                mm = re.fullmatch(r"CH_SPACEDIM\s*(==|>|<|>=|<=|!=)\s*(\d+)", rest)
                if not mm:
                    raise ChfError(f"unsupported preprocessor condition: {line}")
                val = eval(f"{SPACEDIM} {mm.group(1)} {mm.group(2)}")
                if kind == "if":
                    stack.append([val, val])
                else:
                    stack[-1][1] = (not stack[-1][0]) and val
                    stack[-1][0] = stack[-1][0] or val
            elif kind == "else":
                stack[-1][1] = not stack[-1][0]
                stack[-1][0] = True
            else:
                stack.pop()
            continue
        if not all(s[1] for s in stack):
            continue
        if not line.strip() or line[0] in "Cc*!":
            continue
        if len(line) > 5 and line[:5].strip() == "" and line[5] not in " 0":
            if not out:
                raise ChfError("continuation without a statement")
            out[-1] += " " + line[6:].strip()
        elif out and out[-1].count("[") > out[-1].count("]"):
            out[-1] += " " + line.strip()          # a CHF_ macro's [...] may span lines without continuation marks
        else:
            out.append(line.strip())
    return out


def split_top(s, sep=";"):
    """split at separators that are not inside () or []"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def find_macro(s, name):
    """first NAME[...] in s: (start, end, inner) with balanced brackets, or None"""
    m = re.search(r"\b" + name + r"\s*\[", s)
    if not m:
        return None
    depth, i = 1, m.end()
    while depth:
        if i >= len(s):
            raise ChfError(f"unbalanced {name}[ in: {s}")
        depth += {"[": 1, "]": -1}.get(s[i], 0)
        i += 1
    return m.start(), i, s[m.end():i - 1]


def expand_expr_macros(s):
    """the macros that may appear inside expressions / declarations"""
    while True:
        for name in ("CHF_IX", "CHF_AUTOIX", "CHF_OFFSETIX", "CHF_DDECL", "CHF_AUTODECL", "CHF_NCOMP", "CHF_LBOUND", "CHF_UBOUND",
                     "CHF_DTERM"):
            hit = find_macro(s, name)
            if not hit:
                continue
            a, b, inner = hit
            p = split_top(inner)
            if name in ("CHF_IX", "CHF_DDECL"):
                rep = ",".join(p[:SPACEDIM])
            elif name in ("CHF_AUTOIX", "CHF_AUTODECL"):
                rep = ",".join(f"{p[0]}{d}" for d in range(SPACEDIM))
            elif name == "CHF_OFFSETIX":
                mm = re.fullmatch(r"([+-])\s*(?:(\d+)\s*\*\s*)?(\w+)", p[1])
                if not mm:
                    raise ChfError(f"unsupported CHF_OFFSETIX offset: {p[1]}")
                sign, mult, off = mm.group(1), mm.group(2), mm.group(3)
                rep = ",".join(f"{p[0]}{d}{sign}{(mult + '*') if mult else ''}{off}{d}" for d in range(SPACEDIM))
            elif name == "CHF_NCOMP":
                rep = f"(*{p[0]}_ncomp)"
            elif name == "CHF_LBOUND":
                rep = f"(*{p[0]}_lo{int(p[1])})"
            elif name == "CHF_UBOUND":
                rep = f"(*{p[0]}_hi{int(p[1])})"
            else:   # CHF_DTERM inside an expression: the first SPACEDIM parts, concatenated
                rep = " ".join(p[:SPACEDIM])
            s = s[:a] + rep + s[b:]
            break
        else:
            break
    # D_TERM(a, b, c) and CHF_ID(a, b)
    while True:
        m = re.search(r"\bD_TERM\s*\(", s)
        if not m:
            break
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(s[i], 0)
            i += 1
        s = s[:m.start()] + "(" + " ".join(split_top(s[m.end():i - 1], ",")[:SPACEDIM]) + ")" + s[i:]
    s = re.sub(r"\bCHF_ID\s*\(", "chf_id(", s)
    return s


def fortran_ops(s):
    for f, c in ((".ne.", "!="), (".eq.", "=="), (".lt.", "<"), (".le.", "<="), (".gt.", ">"), (".ge.", ">="), (".and.", "&&"),
                 (".or.", "||"), (".not.", "!")):
        s = re.sub(re.escape(f), f" {c} ", s, flags=re.I)
    s = s.replace("CH_SPACEDIM", str(SPACEDIM))
    if "**" in s:
        raise ChfError(f"power operator not supported: {s}")
    return s


ARG = {"CHF_FRA": ("double *", True, True), "CHF_CONST_FRA": ("const double *", True, True), "CHF_FRA1": ("double *", True, False),
       "CHF_CONST_FRA1": ("const double *", True, False)}


def translate_subroutine(header, body):
    m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", header, flags=re.I)
    if not m:
        raise ChfError(f"cannot parse: {header}")
    name, args = m.group(1), split_top(m.group(2), ",")
    params, prologue = [], []
    for a in args:
        mm = re.fullmatch(r"(CHF_\w+)\s*\[\s*(\w+)\s*\]", a)
        if not mm:
            raise ChfError(f"unsupported argument: {a}")
        kind, v = mm.group(1), mm.group(2)
        bounds = [f"const int *{v}_{w}{d}" for w in ("lo", "hi") for d in range(SPACEDIM)]
        if kind in ARG:
            ptr, _, has_comp = ARG[kind]
            params += [f"{ptr}{v}_data"] + bounds + ([f"const int *{v}_ncomp"] if has_comp else [])
            idx = (f"({v}_data[(i - *{v}_lo0) + (long)(*{v}_hi0 - *{v}_lo0 + 1) * ((j - *{v}_lo1) + (long)(*{v}_hi1 - *{v}_lo1 + 1) * "
                   f"((k - *{v}_lo2)" + (f" + (long)(*{v}_hi2 - *{v}_lo2 + 1) * n" if has_comp else "") + "))])")
            ref = "const double &" if "const" in ptr else "double &"
            sig = "int i, int j, int k, int n" if has_comp else "int i, int j, int k"
            prologue.append(f"  auto {v} = [=]({sig}) -> {ref} {{ return {idx}; }};")
        elif kind == "CHF_BOX":
            params += bounds
        elif kind == "CHF_CONST_REAL":
            params.append(f"const double *{v}_p")
            prologue.append(f"  const double {v} = *{v}_p;")
        elif kind == "CHF_CONST_INT":
            params.append(f"const int *{v}_p")
            prologue.append(f"  const int {v} = *{v}_p;")
        else:
            raise ChfError(f"unsupported argument kind: {kind}")
    out = [f'extern "C" void {name.lower()}_({", ".join(params)}) {{'] + prologue
    depth = 1
    for line in body:
        low = line.lower()
        if re.fullmatch(r"return", low):
            out.append("  " * depth + "return;")
            continue
        if re.fullmatch(r"end", low):
            break
        hit = re.match(r"(CHF_MULTIDO|CHF_AUTOMULTIDO)\s*\[(.*)\]\s*$", line)
        if hit:
            p = split_top(hit.group(2))
            box = p[0]
            ivs = p[1:1 + SPACEDIM] if hit.group(1) == "CHF_MULTIDO" else [f"{p[1]}{d}" for d in range(SPACEDIM)]
            for d in reversed(range(SPACEDIM)):
                out.append("  " * depth + f"for ({ivs[d]} = *{box}_lo{d}; {ivs[d]} <= *{box}_hi{d}; {ivs[d]}++) {{")
                depth += 1
            continue
        if low == "chf_enddo":
            for _ in range(SPACEDIM):
                depth -= 1
                out.append("  " * depth + "}")
            continue
        hit = find_macro(line, "CHF_DTERM")
        if hit and hit[0] == 0 and hit[1] == len(line) and all("=" in q for q in split_top(hit[2])[:SPACEDIM]):
            for q in split_top(hit[2])[:SPACEDIM]:          # a block of statements, one per direction
                out.append("  " * depth + fortran_ops(expand_expr_macros(q)) + ";")
            continue
        s = fortran_ops(expand_expr_macros(line))
        low = s.lower()
        mm = re.match(r"(real_t|integer)\s+(.*)$", s, flags=re.I)
        if mm:
            out.append("  " * depth + ("double " if mm.group(1).lower() == "real_t" else "int ") + mm.group(2) + ";")
            continue
        mm = re.match(r"do\s+(\w+)\s*=\s*(.*)$", s, flags=re.I)
        if mm:
            v, rng = mm.group(1), split_top(mm.group(2), ",")
            if len(rng) not in (2, 3):
                raise ChfError(f"cannot parse do loop: {line}")
            step = rng[2] if len(rng) == 3 else "1"
            if not re.fullmatch(r"\d+", step):
                raise ChfError(f"only positive literal do-steps are supported: {line}")
            out.append("  " * depth + f"for ({v} = {rng[0]}; {v} <= {rng[1]}; {v} += {step}) {{")
            depth += 1
            continue
        if low in ("enddo", "end do", "endif", "end if"):
            depth -= 1
            out.append("  " * depth + "}")
            continue
        mm = re.match(r"if\s*\((.*)\)\s*then$", s, flags=re.I)
        if mm:
            out.append("  " * depth + f"if ({mm.group(1)}) {{")
            depth += 1
            continue
        if re.fullmatch(r"call\s+maydayerror\s*\(\s*\)", low):
            out.append("  " * depth + "abort();")
            continue
        if re.match(r"[\w]+(\s*\(.*\))?\s*=[^=]", s):
            out.append("  " * depth + s + ";")
            continue
        raise ChfError(f"unsupported statement: {line}")
    if depth != 1:
        raise ChfError(f"unbalanced blocks in {name}")
    out.append("}")
    return "\n".join(out)


def translate(text, source="<stdin>"):
    lines = preprocess(text)
    units, i = [], 0
    while i < len(lines):
        if not re.match(r"subroutine\b", lines[i], flags=re.I):
            raise ChfError(f"statement outside a subroutine: {lines[i]}")
        j = i + 1
        while j < len(lines) and not re.fullmatch(r"end", lines[j], flags=re.I):
            j += 1
        if j == len(lines):
            raise ChfError("subroutine without end")
        units.append(translate_subroutine(lines[i], lines[i + 1:j + 1]))
        i = j + 1
    head = (f"// GENERATED by oracle/chf2c.py from {source} -- do not edit, do not commit.\n"
            "#include <cmath>\n#include <cstdlib>\nusing std::abs;\n" + CONSTANTS + "\n"
            "static inline int mod(int a, int b) { return a % b; }          // Fortran MOD: sign of the dividend, like C\n"
            "static inline int chf_id(int a, int b) { return a == b ? 1 : 0; }\n\n")
    return head + "\n\n".join(units) + "\n"


if __name__ == "__main__":
    path = sys.argv[1]
    with open(path) as f:
        sys.stdout.write(translate(f.read(), path))
