"""Host-side mirror of the reference's inertia-correcting regularisation:

  RegularizationStrategy            LinearSystems/RegularizationStrategy.cpp:13-21   (is_inertia_correct)
  IpoptRegularization               LinearSystems/IpoptRegularization.cpp:30-92      (REGULARIZATION_STRATEGY = 0)
  FriedlanderOrbanRegularization    LinearSystems/FriedlanderOrbanRegularization.cpp:28-52 (REGULARIZATION_STRATEGY = 1)
  factorize_with_correct_inertia    LinearSystems/LinearSystem.C:296-325, called per node from
                                    DistributedLeafLinearSystem::factor2 (:74-86) and
                                    DistributedRootLinearSystem::factorizeKKT (:1459)

Pure control logic (scalars only); the factorisations and the inertia come from the CUDA path through the C ABI
(pipsgpu_leaf_get_inertia counts the signs of the COMPUTED pivots, so a wrong-signed or vanishing pivot is visible).
The C++ rendition a PIPS-IPM++ build uses is the reference's own LinearSystem::factorize_with_correct_inertia, unchanged:
the shim solvers (include/pipsgpu_shim.hpp) report reports_inertia() == true and the true inertia."""
from __future__ import annotations


class RegularizationStrategy:
    def __init__(self, positive_expected: int, negative_expected: int):
        self.positive_expected, self.negative_expected = positive_expected, negative_expected
        self.new_factorization = True

    def is_inertia_correct(self, inertia) -> bool:
        pos, neg, zero = inertia
        return pos == self.positive_expected and neg == self.negative_expected and zero == 0

    def notify_new_step(self):
        self.new_factorization = True


class IpoptRegularization(RegularizationStrategy):
    barrier_exponent_dual = 0.25
    primal_initial = 1e-4
    primal_decrease_factor = 1.0 / 3.0
    primal_increase_factor_initial = 100.0
    primal_increase_factor = 8.0

    def __init__(self, positive_expected, negative_expected, primal_min=1e-20, primal_max=1e40):
        super().__init__(positive_expected, negative_expected)
        self.primal_min, self.primal_max = primal_min, primal_max
        self.primal_last, self.primal_current = 0.0, -1.0
        self.dual_eq_current = self.dual_ineq_current = 0.0

    def notify_new_step(self):
        super().notify_new_step()
        if self.primal_current != -1.0:
            self.primal_last = self.primal_current

    def get_default_regularization(self):
        return 0.0, 0.0, 0.0

    def get_regularization_parameters(self, inertia, barrier_parameter):
        if self.new_factorization:
            self.new_factorization = False
            zero = inertia[2]
            self.dual_eq_current = barrier_parameter ** self.barrier_exponent_dual if zero else 1e-4
            if self.primal_last == 0.0:
                self.primal_current = self.primal_initial
            else:
                self.primal_current = max(self.primal_min, self.primal_decrease_factor * self.primal_last)
            self.dual_ineq_current = self.dual_eq_current
        else:
            self.primal_current *= self.primal_increase_factor_initial if self.primal_last == 0.0 else self.primal_increase_factor
            if self.primal_current > self.primal_max:
                raise RuntimeError("ERROR : cannot factorize matrix after excessive error correction")
        return self.primal_current, self.dual_eq_current, self.dual_ineq_current


class FriedlanderOrbanRegularization(RegularizationStrategy):
    increase_factor = 100.0
    decrease_factor = 0.1

    def __init__(self, positive_expected, negative_expected, initial=(1.0, 1.0, 1.0), primal_min=1e-10, dual_min=1e-10):
        super().__init__(positive_expected, negative_expected)
        self.primal_min = primal_min
        self.dual_min = primal_min      # the reference reads ..._PRIMAL_MIN for both (FriedlanderOrbanRegularization.cpp:13)
        self.cur = [v / self.decrease_factor for v in initial]

    def get_default_regularization(self):
        return self.get_regularization_parameters(None, None)

    def get_regularization_parameters(self, inertia=None, barrier_parameter=None):
        if self.new_factorization:
            self.new_factorization = False
            self.cur[0] = max(self.primal_min, self.cur[0] * self.decrease_factor)
            self.cur[1] = max(self.dual_min, self.cur[1] * self.decrease_factor)
            self.cur[2] = max(self.dual_min, self.cur[2] * self.decrease_factor)
        else:
            self.cur = [v * self.increase_factor for v in self.cur]
        return tuple(self.cur)


def factorize_with_correct_inertia(strategy, add_regularization, matrix_changed, get_inertia, barrier_parameter,
                                   max_tries=60):
    """LinearSystem::factorize_with_correct_inertia (LS/LinearSystem.C:296-325) for ONE node of the tree.
    add_regularization(dP, dDy, dDz) adds increments to the node's KKT diagonal, matrix_changed() refactorises,
    get_inertia() -> (pos, neg, zero). Returns (final regularisation triple, number of factorisations)."""
    strategy.notify_new_step()
    last = strategy.get_default_regularization()
    add_regularization(*last)
    matrix_changed()
    nfact = 1
    while not strategy.is_inertia_correct(get_inertia()):
        if nfact >= max_tries:
            raise RuntimeError("factorize_with_correct_inertia: inertia still wrong after %d factorisations" % nfact)
        cur = strategy.get_regularization_parameters(get_inertia(), barrier_parameter)
        assert all(c >= l for c, l in zip(cur, last))
        add_regularization(*(c - l for c, l in zip(cur, last)))
        last = cur
        matrix_changed()
        nfact += 1
    return last, nfact
