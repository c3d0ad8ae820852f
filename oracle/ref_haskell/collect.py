"""Output of Golden.hs -> tests/golden/reference.json (see README.md).
    python oracle/ref_haskell/collect.py _ref/ref_output.txt"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def pairs(tokens):
    x = [float(t) for t in tokens]
    return [[x[i], x[i + 1]] for i in range(0, len(x), 2)]


def main(path):
    ref = {"what": "amplitudes computed by the unmodified reference (qubitrot/qubism, stack lts-12.4) on the cases of golden.json",
           "gate_vectors": {}, "programs": {}}
    for line in open(path):
        w = line.split()
        if not w:
            continue
        if w[0] == "GATE":
            ref["gate_vectors"][w[1]] = pairs(w[2:])
        elif w[0] == "PROG":
            name, run = w[1], w[2]
            rec = {"states": {}, "cregs": {}, "error": None}
            body = line.split(None, 3)[3].strip()
            if body.startswith(("PARSE-ERROR", "RUNTIME-ERROR")):
                rec["error"] = body
            else:
                for part in body.split(" ; "):
                    t = part.split()
                    if t[0] == "SV":
                        rec["states"][t[1]] = pairs(t[2:])
                    elif t[0] == "CREG":
                        rec["cregs"][t[1]] = [int(ch) for ch in t[2]] if len(t) > 2 else []
            ref["programs"].setdefault(name, {})[run] = rec
    out = os.path.join(ROOT, "tests", "golden", "reference.json")
    json.dump(ref, open(out, "w"))
    print("wrote", out, len(ref["gate_vectors"]), "gate vectors,", sum(len(v) for v in ref["programs"].values()), "program runs")


if __name__ == "__main__":
    main(sys.argv[1])
