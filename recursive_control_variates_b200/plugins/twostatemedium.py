"""plugins/twostatemedium.py of the reference: a medium holding theta_{i-1} (`old`),
theta_i (`new`) and the optimiser-facing `incoming` parameters.

The three Python-side media only mirror the reference's bookkeeping; the grids
the kernels read are the context's packed two-state arrays, rotated by
rcv_medium_rotate when the scene parameters are updated
(mitsuba_compat.SceneParameters.update -> Scene.push_medium)."""
from __future__ import annotations


class TwoStateMedium:
    def __init__(self, props):
        self.old = props["old"]            # twostatemedium.py:11-13
        self.new = props["new"]
        self.incoming = props["incoming"]

    def is_twostate(self) -> bool:
        return True

    def incoming_medium(self):
        return self.incoming

    def volume_boxes(self):
        return self.old.volume_boxes() + self.new.volume_boxes() + self.incoming.volume_boxes()

    # everything except the two-state sampling delegates to `new` (twostatemedium.py:117-184)
    def phase_function(self):
        return self.new.phase

    def use_emitter_sampling(self):
        return self.new.sample_emitters

    def is_homogeneous(self):
        return False

    # majorant supergrid (twostatemedium.py:141-172): both states share the factor
    def majorant_resolution_factor(self) -> int:
        return self.new.majorant_resolution_factor()

    def set_majorant_resolution_factor(self, factor: int):
        self.old.set_majorant_resolution_factor(factor)      # :150-154
        self.new.set_majorant_resolution_factor(factor)

    def has_majorant_grid(self) -> bool:
        return self.new.has_majorant_grid()

    def traverse(self):
        """Only `incoming` is exposed to the optimiser (twostatemedium.py:186-187)."""
        return self.incoming.parameters()

    def parameters_changed(self, keys=None):
        """old <- new, new <- incoming (twostatemedium.py:189-200).  References only: the
        device copy and the majorant update happen in rcv_medium_rotate."""
        self.old.sigma_t.data, self.old.albedo.data = self.new.sigma_t.data, self.new.albedo.data
        self.new.sigma_t.data, self.new.albedo.data = self.incoming.sigma_t.data, self.incoming.albedo.data

    def to_string(self):
        return "TwoStateMedium[old, new, incoming]"
