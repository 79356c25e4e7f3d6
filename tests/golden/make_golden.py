"""Bundle the reference's golden vectors for the LP/MIP path into one fixture file.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py

Inputs : /root/reference/test/test-sanity/*.json  (47 models, each with an `expects` block;
         compared by src/solver.integration.test.ts:60-100,144-166) and the README known answers.
Output : tests/golden/sanity_fixtures.json.gz -- {"fixtures": [{"file", "model", "expects"}...],
         "readme": [...]}.  Key order inside every model is preserved (it decides row/column
         order, model.ts:288,334).  The GPU box has no /root/reference: tests read only the bundle.
"""
import glob
import gzip
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sanity_fixtures.json.gz")

README = [
    {   # README.md:57-73
        "file": "README Berlin Airlift",
        "model": {
            "optimize": "capacity", "opType": "max",
            "constraints": {"plane": {"max": 44}, "person": {"max": 512}, "cost": {"max": 300000}},
            "variables": {
                "brit": {"capacity": 20000, "plane": 1, "person": 8, "cost": 5000},
                "yank": {"capacity": 30000, "plane": 1, "person": 16, "cost": 9000}},
        },
        "expects": {"feasible": True, "result": 1080000, "brit": 24, "yank": 20},
    },
    # README.md:128-147 (tables/dressers, "result: 14400, table: 8, dresser: 3") is NOT included:
    # the printed answer is not optimal for the printed model (table 4, dresser 9 is feasible
    # and gives 19200), so it is a documentation slip, not a known answer.
    {   # README.md:30-38
        "file": "README quick example",
        "model": {
            "optimize": "profit", "opType": "max",
            "constraints": {"capacity": {"max": 100}},
            "variables": {"x": {"capacity": 10, "profit": 5}},
        },
        "expects": {"feasible": True, "result": 50, "x": 10},
    },
    {   # src/solver.integration.test.ts:115-141
        "file": "integration widget",
        "model": {
            "optimize": "profit", "opType": "max",
            "constraints": {"capacity": {"max": 5}},
            "variables": {"widget": {"capacity": 1, "profit": 1}},
            "ints": {"widget": 1},
        },
        "expects": {"feasible": True, "widget": 5, "result": 5},
    },
]


def main():
    fixtures = []
    for path in sorted(glob.glob(os.path.join(REF, "test", "test-sanity", "*.json"))):
        with open(path) as f:
            model = json.load(f)
        expects = model.pop("expects")
        fixtures.append({"file": os.path.basename(path), "model": model, "expects": expects})
    blob = json.dumps({"source": "JWally/jsLPSolver@7a15082 test/test-sanity", "fixtures": fixtures,
                       "readme": README}, separators=(",", ":")).encode()
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(blob)
    print(f"wrote {OUT}: {len(fixtures)} fixtures, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
