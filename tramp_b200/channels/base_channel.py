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
