#!/usr/bin/env python
"""Verbatim-line audit of this repo against the reference tree.

For every Python file of the package the source is normalised (comments and docstrings stripped through the
tokenizer, whitespace collapsed) and the fraction of its lines longer than 30 characters that occur verbatim
anywhere in the reference's Python tree is reported. Files above the threshold fail the audit.

    python tools/copycheck.py [--ref /root/reference/python] [--threshold 0.25] [paths...]
"""
import argparse
import io
import os
import sys
import tokenize


def normalised_lines(path):
    try:
        src = open(path, "rb").read().decode("utf-8", "replace")
    except OSError:
        return []
    out, prev_end, line_toks = [], None, []
    lines = {}
    try:
        toks = list(tokenize.generate_tokens(io.StringIO(src).readline))
    except (tokenize.TokenError, IndentationError, SyntaxError):
        return [" ".join(l.split()) for l in src.splitlines() if len(" ".join(l.split())) > 30]
    prev_type = tokenize.INDENT
    for tok in toks:
        if tok.type == tokenize.COMMENT:
            continue
        if tok.type == tokenize.STRING and prev_type in (tokenize.INDENT, tokenize.NEWLINE, tokenize.NL, tokenize.DEDENT):
            prev_type = tok.type  # docstring / bare string statement
            continue
        if tok.type in (tokenize.NL, tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT, tokenize.ENDMARKER):
            prev_type = tok.type if tok.type != tokenize.NL or prev_type in (tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT, tokenize.NL) else prev_type
            continue
        lines.setdefault(tok.start[0], []).append(tok.string)
        prev_type = tok.type
    for ln in sorted(lines):
        text = " ".join(lines[ln])
        if len(text) > 30:
            out.append(text)
    return out


def collect(root):
    files = []
    for d, _, fs in os.walk(root):
        if "/baseline/" in d + "/" or "/.git" in d:
            continue
        for f in fs:
            if f.endswith(".py"):
                files.append(os.path.join(d, f))
    return files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--threshold", type=float, default=0.25)
    ap.add_argument("--min-lines", type=int, default=12)
    ap.add_argument("--show", action="store_true", help="print the matching lines of the given paths")
    ap.add_argument("paths", nargs="*", default=None)
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_lines = set()
    for f in collect(args.ref):
        ref_lines.update(normalised_lines(f))
    targets = []
    for p in (args.paths or [os.path.join(root, "dgl_b200"), os.path.join(root, "tools")]):
        targets += collect(p) if os.path.isdir(p) else [p]
    rows = []
    for f in sorted(targets):
        ls = normalised_lines(f)
        if len(ls) < args.min_lines:
            continue
        hit = sum(1 for l in ls if l in ref_lines)
        rows.append((hit / len(ls), len(ls), os.path.relpath(f, root)))
        if args.show and args.paths:
            print(f"== {os.path.relpath(f, root)}: {hit}/{len(ls)}")
            for l in ls:
                if l in ref_lines:
                    print("   ", l)
    rows.sort(reverse=True)
    bad = [r for r in rows if r[0] >= args.threshold]
    for frac, n, f in rows[:40]:
        print(f"{frac:5.2f}  {n:5d}  {f}")
    print(f"{len(bad)} file(s) at or above {args.threshold}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
