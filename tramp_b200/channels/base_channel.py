from ..base import Factor


class Channel(Factor):
    """reference channels/base_channel.py:5-17."""
    n_next = 1
    n_prev = 1

    def compute_forward_message(self, az, bz, ax, bx):
        rx, vx = self.compute_forward_posterior(az, bz, ax, bx)
        return self.compute_ab_new(rx, vx, ax, bx)

    def compute_backward_message(self, az, bz, ax, bx):
        rz, vz = self.compute_backward_posterior(az, bz, ax, bx)
        return self.compute_ab_new(rz, vz, az, bz)

    # ---- State Evolution (reference channels/base_channel.py:19-53) ------------
    def compute_forward_state_evolution(self, az, ax, tau_z):
        vx = self.compute_forward_error(az, ax, tau_z)
        return self.compute_a_new(vx, ax)

    def compute_backward_state_evolution(self, az, ax, tau_z):
        vz = self.compute_backward_error(az, ax, tau_z)
        return self.compute_a_new(vz, az)

    def compute_forward_overlap(self, az, ax, tau_z):
        return self.second_moment(tau_z) - self.compute_forward_error(az, ax, tau_z)

    def compute_backward_overlap(self, az, ax, tau_z):
        return tau_z - self.compute_backward_error(az, ax, tau_z)

    def get_alpha(self):
        return getattr(self, "alpha", 1)
