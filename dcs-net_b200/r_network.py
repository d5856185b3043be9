"""Drop-in twin of the model classes of the reference's r_network.py (dr / drs): RealChannelAttention,
RealSpatialAttention and R_NETWORK(config, hparams, seed) — same constructor signatures, attribute names and registration
order, hence the same 158 state_dict keys / shapes and, for a given seed, bit-identical random-init weights (pinned by
tests/test_rnet_oracle.py against tests/golden/rnet_*.pt, which were produced by the reference's own r_network.py).

ROUND-1 STATUS: the parameter containers, the CPU oracle (oracle/rnet_oracle.py) and its reference-generated golden
vectors exist; the sm_100a kernels of the real path do not yet (real convs reuse the implicit-GEMM kernels with real
packing, but the max-pool channel attention, the real spatial attention, the 256 -> 128 real LSTM and the
sigmoid / magnitude-mask tail need their own kernels).  `forward` therefore raises — there is no ATen / CPU fallback.
"""
import torch

from .c_network import _Base, _seed_everything


class RealChannelAttention(torch.nn.Module):
    """r_network.py:8-25 (only the max-pool branch reaches the output)."""

    def __init__(self, no_channels, reduction_ratio):
        super().__init__()
        self.avg_pool = torch.nn.AdaptiveAvgPool2d(1)
        self.max_pool = torch.nn.AdaptiveMaxPool2d(1, return_indices=False)
        self.fc = torch.nn.Sequential(torch.nn.Conv2d(no_channels, max(no_channels // reduction_ratio, 1), 1, bias=False),
                                      torch.nn.ReLU(),
                                      torch.nn.Conv2d(max(no_channels // reduction_ratio, 1), no_channels, 1, bias=False))
        self.sigmoid = torch.nn.Sigmoid()

    def forward(self, x):
        raise NotImplementedError("dcsnet_b200: the real (dr / drs) path has no sm_100a kernels yet (SURVEY 8f rank 1); no CPU fallback")


class RealSpatialAttention(torch.nn.Module):
    """r_network.py:28-40."""

    def __init__(self, kernel_size):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = torch.nn.Sigmoid()

    def forward(self, x):
        raise NotImplementedError("dcsnet_b200: the real (dr / drs) path has no sm_100a kernels yet (SURVEY 8f rank 1); no CPU fallback")


class R_NETWORK(_Base):
    """r_network.py:43-173."""

    def __init__(self, config, hparams, seed):
        super().__init__()
        _seed_everything(seed)
        self.config = config
        self.hparams.update(hparams)
        self.save_hyperparameters(self.hparams)
        hp = self.hparams
        self.encoder = torch.nn.ModuleList()
        self.decoder = torch.nn.ModuleList()
        self.decoder_attention = torch.nn.ModuleList()
        self.skip_attention = torch.nn.ModuleList()
        n_layers, ch = hp['no_of_layers'], hp['channels']
        self.initial_batchnorm = torch.nn.BatchNorm2d(ch[0])
        for i in range(n_layers):
            self.encoder.append(torch.nn.Sequential(
                torch.nn.Conv2d(in_channels=1 if i == 0 else ch[i], out_channels=ch[i + 1], kernel_size=config.kernel_sizeE[i],
                                stride=config.strideE[i], padding=config.paddingE[i]),
                torch.nn.BatchNorm2d(ch[i + 1]),
                config.RactivationE()))
        self.lstm = torch.nn.LSTM(input_size=ch[5], hidden_size=ch[4], num_layers=hp['lstm_layers'],
                                  bidirectional=hp['lstm_bidir'], batch_first=True)
        self.fc = torch.nn.Linear(ch[5], ch[5])
        self.dropout_conv = torch.nn.Dropout(hp['dropout_conv'])
        self.dropout_fc = torch.nn.Dropout(hp['dropout_fc'])
        for i in range(n_layers):
            in_channels = ch[n_layers - i]
            out_channels = max(ch[n_layers - 1 - i], 1)
            convt = torch.nn.ConvTranspose2d(in_channels + in_channels, out_channels, kernel_size=config.kernel_sizeD[i],
                                             stride=config.strideD, padding=config.paddingD[i])
            if i == n_layers - 1:
                self.decoder.append(convt)
            else:
                self.decoder.append(torch.nn.Sequential(convt, torch.nn.BatchNorm2d(ch[n_layers - 1 - i]), config.RactivationD()))
            self.skip_attention.append(RealChannelAttention(in_channels, hp['channel_attention_reduction_ratio']))
            self.skip_attention.append(RealSpatialAttention(hp['spatial_attention_kernel_size']))
            self.decoder_attention.append(RealChannelAttention(out_channels, hp['channel_attention_reduction_ratio']))
            self.decoder_attention.append(RealSpatialAttention(hp['spatial_attention_kernel_size']))
        self.weights_init()

    def weights_init(self):
        init = self.hparams['initialisation_distribution']
        for m in self.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
                init(m.weight)

    def forward(self, x):
        raise NotImplementedError("dcsnet_b200.R_NETWORK: the real (dr / drs) forward has no sm_100a kernels yet "
                                  "(SURVEY 8f rank 1; oracle: oracle/rnet_oracle.py); there is no CPU / ATen fallback")
