"""Drop-in for the GNAN classes of the reference's models.py (what main.py imports, main.py:2,77-90):
`TensorGNAN` (models.py:303-384), `GNAN` (models.py:387-481, identical math to GNAN.py's, kw `num_layers`).

Differences from GNAN.py's TensorGNAN that are honoured here (models.py:320-321,366-370):
  * rho has ONE output unless `rho_per_feature` (then out_channels);
  * `normalize_rho` divides rho's OUTPUT by the normalisation matrix;
  * graph task with `readout_n_layers > 0` (a NAM over the per-feature pooled values, :348-350,380-381) is not on the
    fused path yet and raises NotImplementedError; main.py's default is readout_n_layers=0.
"""
import torch

from . import ops
from ._inputs import resolve
from ._stacked import StackedMLP
from .GNAN import GNAN as _GNAN, _Base
from .preprocess import PackedBatch


class TensorGNAN(_Base):
    def __init__(self, in_channels, out_channels, n_layers, hidden_channels=None, bias=True, dropout=0.0,
                 device='cpu', rho_per_feature=False, normalize_rho=True, is_graph_task=False, readout_n_layers=1):
        super().__init__()
        self.device = device
        self.out_channels = out_channels
        self.hidden_channels = hidden_channels
        self.n_layers = n_layers
        self.bias = bias
        self.dropout = dropout
        self.rho_per_feature = rho_per_feature
        self.normalize_rho = normalize_rho
        self.is_graph_task = is_graph_task
        self.readout_n_layers = readout_n_layers
        if is_graph_task and readout_n_layers > 0:
            raise NotImplementedError("readout_n_layers > 0 (NAM readout, models.py:348-350) is not implemented in "
                                      "gnan_b200 yet; main.py's default is 0")
        self.actual_output_dim_f = out_channels
        self.actual_output_dim_rho = out_channels if rho_per_feature else 1
        self.fs = StackedMLP(in_channels, self.actual_output_dim_f, n_layers, hidden_channels, bias, 3, dropout)
        self.rho = StackedMLP(1, self.actual_output_dim_rho, n_layers, hidden_channels, not is_graph_task, 2, single=True)
        self.fs.xavier_normal_(0.01)
        self.rho.xavier_normal_(0.01)

    def forward(self, inputs):
        if isinstance(inputs, PackedBatch):
            return self.forward_packed(inputs)
        x, hd = resolve(inputs, self._device())
        S = self._feature_sums(x)
        T = self._table(ops.rho_table_inputs(hd.nbins, x.device))                    # [nbins,Cr]
        rs = ops.level_rscale(hd.level_counts) if self.normalize_rho else None       # models.py:368-370
        out = ops.aggregate_rows(hd.hop, T, S, rscale=rs)
        if self.is_graph_task:
            out = out.sum(dim=0).view(1, -1).T
        return out


    def forward_packed(self, pk):
        """Many graphs in one call (extension; see gnan_b200.GNAN.TensorGNAN.forward_packed): [B,C] for graph tasks."""
        dev = self._device()
        if pk.x.device != dev:
            pk = pk.to(dev)
        S = self._feature_sums(pk.x.float().contiguous())
        T = self._table(ops.rho_table_inputs(pk.nbins, dev))
        rs = ops.level_rscale(pk.level_counts) if self.normalize_rho else None
        return ops.aggregate_blockdiag(pk.hop, pk.hop_off, pk.node_off, T, S, rscale=rs, reduce_graph=self.is_graph_task)


class GNAN(_GNAN):
    def __init__(self, in_channels, out_channels, num_layers, hidden_channels=None, bias=True, dropout=0.0,
                 device='cpu', normalize_rho=True, rho_per_feature=False):
        super().__init__(in_channels, out_channels, n_layers=num_layers, hidden_channels=hidden_channels, bias=bias,
                         dropout=dropout, device=device, normalize_rho=normalize_rho, rho_per_feature=rho_per_feature)
