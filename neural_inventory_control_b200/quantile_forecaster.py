"""Quantile forecaster used by the quantile policies (reference: quantile_forecaster.py:5-161).

A plain MLP `net` (state_dict keys `net.<2i>.{weight,bias}`, so the reference's shipped checkpoint
`quantile_forecasters/1700580865.pt` loads unchanged) that predicts, for every (sample, store), the quantiles
q = 0.05 .. 0.95 of cumulative demand over each lead time; `get_quantile` inverts a requested quantile into a
base-stock level by linear interpolation between the two neighbouring predicted quantiles. Runs as ordinary
torch ops inside the generic per-step path (the forecaster is frozen: no gradient work, nothing to fuse).
"""
import numpy as np
import torch
from torch import nn


class FullyConnectedForecaster(nn.Module):
    def __init__(self, neurons_per_hidden_layer, lead_times, qs=np.arange(0.05, 1, 0.05), activation_function=None,
                 device=None):
        super().__init__()
        self.qs = np.asarray(qs).round(2)
        self.qs_dict = {round(float(q), 2): i for i, q in enumerate(qs)}
        self.lead_times = torch.tensor(lead_times).int()
        self.min_lead_time = int(min(lead_times))
        self.lead_times_dict = {int(lt): i for i, lt in enumerate(lead_times)}
        self.activation_function = activation_function if activation_function is not None else nn.ELU()
        layers = []
        for width in neurons_per_hidden_layer:
            layers += [nn.LazyLinear(width), self.activation_function]
        layers.append(nn.LazyLinear(len(self.qs) * len(lead_times)))
        self.layers = layers
        self.net = nn.Sequential(*layers)
        # probability grid with the extrapolated end points 0 and 1; a buffer outside the state_dict so that it follows
        # .to(device) (the reference pins it to "cuda if available" at construction time)
        self.register_buffer("prob_points", torch.tensor([0.0] + [float(q) for q in self.qs] + [1.0]), persistent=False)

    def forward(self, x):
        y = torch.clip(self.net(x), min=0)
        return y.reshape(*y.shape[:-1], len(self.qs), len(self.lead_times))

    def create_0_1_quantiles(self, x):
        """Linear extrapolation of the first / last predicted quantile to probability 0 / 1 (dim 2 = quantile)."""
        first = 2 * x[:, :, :1] - x[:, :, 1:2]
        last = 2 * x[:, :, -1:] - x[:, :, -2:-1]
        return torch.cat((first, x, last), dim=2)

    def retrieve_corresponding_lead_time(self, x, lead_times):
        """Pick the lead-time slice of each (sample, store): lead times are consecutive integers from min_lead_time."""
        idx = (lead_times - self.min_lead_time).to(torch.int64)
        return torch.gather(x, 3, idx[:, :, None, None].expand(-1, -1, x.shape[2], 1)).squeeze(3)

    def _curve(self, x, lead_times):
        return self.create_0_1_quantiles(self.retrieve_corresponding_lead_time(self.forward(x), lead_times))

    def get_quantile(self, x, quantile, lead_times):
        """x [B,S,F] features, quantile [B,S] in (0,1), lead_times [B,S] -> interpolated demand quantile [B,S]."""
        pp = self.prob_points.to(quantile.dtype)
        hi = torch.searchsorted(pp, quantile.contiguous())
        curve = self._curve(x, lead_times)
        below = torch.gather(curve, 2, (hi - 1).unsqueeze(2)).squeeze(2)
        above = torch.gather(curve, 2, hi.unsqueeze(2)).squeeze(2)
        d_prev = quantile - pp[hi - 1]
        d_next = pp[hi] - quantile
        return below + (above - below) * d_prev / (d_prev + d_next)

    def get_implied_percentile(self, x, lead_time_per_sample, inventory_position, allocation=None,
                               zero_out_no_orders=False):
        """Inverse map: the probability level whose predicted quantile equals the inventory position."""
        curve = self._curve(x, lead_time_per_sample)
        pp = self.prob_points.to(curve.dtype)
        n = pp.shape[0]
        hi = torch.clip(torch.searchsorted(curve, inventory_position.unsqueeze(2)), min=1, max=n - 1).squeeze(2)
        p_prev, p_next = pp[hi - 1], pp[hi]
        below = torch.gather(curve, 2, (hi - 1).unsqueeze(2)).squeeze(2)
        above = torch.gather(curve, 2, hi.unsqueeze(2)).squeeze(2)
        d_prev = inventory_position - below
        d_next = above - inventory_position
        pct = p_prev + (p_next - p_prev) * d_prev / (d_prev + d_next)
        if zero_out_no_orders:
            pct = torch.where(allocation == 0, torch.zeros_like(pct), pct)
        return pct
