"""Extract the layer-name patterns the REFERENCE's profiler uses to classify UNet ops (analyze_results.py:20-87,
`determine_op_type`) into a committed fixture.  They are the only in-tree evidence of the UNet's module structure
(SURVEY §8c): the oracle restatement must have a module for every pattern, and vice versa.
Run in the build container only:  python tests/golden/make_layer_patterns.py  -> tests/golden/layer_patterns.json
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/analyze_results.py"


def extract(path=SRC):
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def determine_op_type"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i].startswith((" ", "\t")))
    rows, outer = [], None
    cond = re.compile(r"^(\s*)if (.+):\s*$")
    ret = re.compile(r"^\s*return '([^']+)'")
    pending = None
    for i in range(start, end):
        m = cond.match(lines[i])
        if m:
            indent, expr = len(m.group(1)), m.group(2)
            lits = re.findall(r"'([^']*)'", expr)
            kind = "startswith" if ".startswith(" in expr else ("equals" if "==" in expr else "contains")
            if "args.qnn" in expr:
                continue
            pending = {"kind": kind, "patterns": lits, "line": i + 1, "indent": indent}
            if indent == 4:
                outer = pending
            else:
                pending["within"] = outer["patterns"][0]
            continue
        m = ret.match(lines[i])
        if m and pending is not None:
            rows.append({"kind": pending["kind"], "patterns": pending["patterns"], "op": m.group(1), "line": pending["line"],
                         **({"within": pending["within"]} if "within" in pending else {})})
            pending = None
    return rows


if __name__ == "__main__":
    rows = extract()
    json.dump({"source": "analyze_results.py:determine_op_type (reference, lines %d-%d)" % (rows[0]["line"], rows[-1]["line"]), "rows": rows},
              open(os.path.join(HERE, "layer_patterns.json"), "w"), indent=1)
    print(len(rows), "patterns")
    for r in rows:
        print(r)
