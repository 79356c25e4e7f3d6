"""`solver.Solve(model)` -- mirror of the reference's public entry point (src/main.ts:94-193)
with the LP/MIP path running on the B200.  Same JSON model in, same result shape out."""
from __future__ import annotations

from typing import Any, Optional

from .model import Model, object_keys


class Solver:
    Model = Model

    def __init__(self):
        self.lastSolvedModel: Optional[Model] = None
        self.engine = 0            # JSLP_OPT_ENGINE for every tableau this solver creates
        self.max_spec_batch = 0    # branch-and-cut speculation width (0 = library default)
        self.node_slots = None     # HBM-resident node batch width (None = auto, 0 = one node at a time)
        self.slot_steps = None
        self.options: dict = {}    # JSLP_OPT_* -> value for every tableau this solver creates (tuning / test aids)

    def Solve(self, model: Any, precision: Optional[float] = None, full: bool = False, validate: bool = False):
        if validate:
            raise NotImplementedError("validation.ts stays host-side TypeScript; not part of the GPU path")
        if not model:
            raise ValueError("Solver requires a model to operate on")  # main.ts:110-112
        if isinstance(model, dict):
            opt = model.get("optimize")
            if isinstance(opt, dict) and len(opt) > 1:
                raise NotImplementedError("multi-objective (polyopt.ts) is a caller of Solve; out of scope")
            if model.get("external"):
                raise NotImplementedError("external solvers (src/external) are out of scope")
            options = model.get("options") or {}
            instance = Model(precision).loadJson(model)
            # selectBranchAndCutService (main.ts:62-83): incremental if explicitly asked, else enhanced if a strategy is named
            if options.get("useIncremental") is True or options.get("nodeSelection") or options.get("branching"):
                instance.branchAndCutOptions = {"nodeSelection": options.get("nodeSelection"), "branching": options.get("branching"),
                                                "useIncremental": options.get("useIncremental") is True}
        else:
            instance = model
        instance.tableau.engine = self.engine
        instance.tableau.max_spec_batch = self.max_spec_batch
        instance.tableau.node_slots = self.node_slots
        instance.tableau.slot_steps = self.slot_steps
        instance.tableau.options.update(self.options)
        solution = instance.solve()
        self.lastSolvedModel = instance
        solution.solutionSet = solution.generateSolutionSet()
        if full:
            return solution
        return self._simplified(solution)

    @staticmethod
    def _simplified(solution) -> dict:  # main.ts:173-193
        head = {"feasible": solution.feasible, "result": solution.evaluation, "bounded": solution.bounded}
        if solution._tableau.isIntegralFlag:
            head["isIntegral"] = True
        values = {k: v for k, v in solution.solutionSet.items() if v != 0}
        out = {}
        keys = object_keys(values)
        for k in keys:  # JS objects list integer-like keys first
            if k.isdigit():
                out[k] = values[k]
        out.update(head)
        for k in keys:
            if not k.isdigit():
                out[k] = values[k]
        return out


solver = Solver()


def Solve(model: Any, precision: Optional[float] = None, full: bool = False, validate: bool = False):
    return solver.Solve(model, precision, full, validate)
