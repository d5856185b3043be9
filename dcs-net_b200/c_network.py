"""Drop-in replacement for the model classes of the reference's c_network.py: ComplexLSTM, ComplexChannelAttention,
ComplexSpatialAttention and C_NETWORK(config, hparams, seed) — same constructor signatures, attribute names,
registration order (hence identical state_dict keys AND identical random-init weights for a given seed), same
`forward(x: complex64 (B,F,T)) -> complex64 (B,F,T)` (squeezed, /root/reference/c_network.py:224).

`C_NETWORK.forward` does not execute its sub-modules one by one: in eval mode it runs the fused kernel plan of
engine.ForwardPlan (packed weights, folded BN, concat/upsample folded into the decoder GEMMs).  The sub-module classes
still work stand-alone (each calls the same kernels) so code that pokes at `net.encoder[i]` keeps working.
"""
import random

import numpy as np
import torch

from . import _lib as L
from . import ops, packing
from .complexLayers import *  # noqa: F401,F403  (mirrors c_network.py:5)
from .complexLayers import ComplexConv2d, ComplexConvTranspose2d, ComplexBatchNorm2d, ComplexLinear, ComplexReLU, _PackedCache, _version_key
from .complexFunctions import complex_upsample, complex_relu, to_cl, from_cl  # noqa: F401
from .network_functions import *  # noqa: F401,F403  (mirrors c_network.py:8)
from .network_functions import ComplexAdaptiveAvgPool2d, ComplexAdaptiveMaxPool2d, ComplexSigmoid
from .engine import ForwardPlan, PackedNet

try:  # a real Lightning install (not present in this image) keeps the LightningModule base for train.py/test.py
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    _seed_everything = pl.seed_everything
except Exception:  # pragma: no cover - the image has no pytorch_lightning
    pl = None

    class _HParams(dict):
        """Attribute-style dict like Lightning's AttributeDict: a missing key is an AttributeError (so copy.deepcopy,
        hasattr and getattr-with-default work)."""

        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name) from None

        __setattr__ = dict.__setitem__

    class _Base(torch.nn.Module):
        """Minimal LightningModule surface used by the hot path: `.hparams`, `save_hyperparameters`, `log_dict`."""

        def __init__(self):
            super().__init__()
            object.__setattr__(self, "_hparams_shim", _HParams())

        @property
        def hparams(self):
            return self._hparams_shim

        def save_hyperparameters(self, *a, **k):
            return None

        def log_dict(self, *a, **k):
            return None

        # what Lightning's Trainer provides to the epoch-end hooks; a driver loop (or a test) sets them
        current_epoch = 0
        global_step = 0
        logger = None

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, hparams_file=None, strict=True, **kwargs):
            """The LightningModule call test.py:20-26 makes: `C_NETWORK.load_from_checkpoint(config=, seed=, checkpoint_path=,
            hparams_file=, map_location=)`.  Reads a Lightning `.ckpt` (`state_dict` + `hyper_parameters`, saved under the
            constructor argument name `hparams` by `save_hyperparameters(self.hparams)`, c_network.py:93) and the optional
            `hparams.yaml` (which wins, as in Lightning); entries a yaml python-object tag carries (the
            `initialisation_distribution` function, config.py:33) fall back to `config.hparams`."""
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            if "state_dict" not in ckpt:
                raise KeyError(f"{checkpoint_path}: not a Lightning checkpoint (no 'state_dict')")
            from .config import hparams as defaults
            hp = dict(defaults)
            hp.update({k: v for k, v in dict(ckpt.get("hyper_parameters") or {}).items() if v is not None})
            if hparams_file is not None:
                hp.update({k: v for k, v in _read_hparams_yaml(hparams_file).items() if v is not None})
            kwargs.setdefault("hparams", hp)
            net = cls(**kwargs)
            net.load_state_dict(ckpt["state_dict"], strict=strict)
            return net

    def _read_hparams_yaml(path):
        import yaml

        class _Loader(yaml.SafeLoader):
            pass
        _Loader.add_multi_constructor("tag:yaml.org,2002:python/", lambda loader, suffix, node: None)
        with open(path) as f:
            return dict(yaml.load(f, Loader=_Loader) or {})

    def _seed_everything(seed):
        seed = int(seed)
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        return seed


class _StepMixin:
    """validation_step / test_step / configure_optimizers of the reference's LightningModules (c_network.py:228-239, 263-302,
    337-372; r_network.py likewise) over the GPU evaluation function.  `variant` (dcs | dc | drs | dr) replaces sys.argv[1]
    and defaults to it; `step_metrics` optionally carries {"pesq": fn, "stoi": fn}.  training_step is SURVEY 8f rank 2."""
    _step_dtype = "complex"
    variant = None
    step_metrics = None

    def _two_mask(self):
        from .network_functions import _variant
        return _variant(self, self.variant) in ("dcs", "drs")

    def configure_optimizers(self):
        optimiser = torch.optim.Adam(self.parameters(), lr=self.hparams['lr'], eps=self.hparams['optim_eps'],
                                     weight_decay=self.hparams['optim_weight_decay'], amsgrad=self.hparams['optim_amsgrad'])
        lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimiser, patience=10)
        return {'optimizer': optimiser, 'lr_scheduler': lr_scheduler, 'monitor': 'val_loss' if self._two_mask() else 'speech_loss'}

    def training_step(self, train_batch, batch_idx):
        """c_network.py:243-261: `train_batch_2_loss` in TRAIN mode (batch-statistic BatchNorm with the running-stat update) and
        the metrics dict, on the GPU through train_engine.TrainStep (complex variants; dropout by dcs_dropout).  Returns the loss as
        a device scalar WITHOUT a grad_fn: there is no autograd graph — the backward pass is `self.train_step.backward()` (hand-written
        kernels filling every parameter's .grad) and the update `self.train_step.optimizer_step()` (global-norm clip + Adam-amsgrad),
        which the train.py shim calls where Lightning would call `loss.backward()` / `optimizer.step()`."""
        from .network_functions import _variant
        variant = _variant(self, self.variant)
        if self._step_dtype != "complex":
            raise NotImplementedError("dcsnet_b200: the real path's training step is not built (SURVEY 8f rank 2)")
        from .train_engine import TrainStep
        if getattr(self, "train_step", None) is None or self.train_step.variant != variant:
            object.__setattr__(self, "train_step", TrainStep(self, variant, speech_alpha=self.hparams["speech_alpha"],
                                                             atan2_eps=self.hparams["atan2_eps"]))
        noise_data, noisy_data, clean_data = train_batch[0], train_batch[1], train_batch[2]
        out = self.train_step.forward(noise_data, noisy_data, clean_data)
        if variant == "dcs":
            metrics = {'train_loss': out["train_loss"].detach(), 'noise_loss': out["noise_loss"].detach(), 'speech_loss': out["speech_loss"].detach()}
            loss = out["train_loss"]
        else:
            metrics = {'speech_loss': out["speech_loss"].detach()}
            loss = out["speech_loss"]
        self.log_dict(metrics, on_epoch=True)
        if torch.any(torch.isnan(loss)):
            print("found NaN in C train loss!")
            return None
        return loss

    @staticmethod
    def _audio_dict(clean, predict_clean, noise, noisy, predict_noise=None):
        out = {"clean": clean.cpu().numpy(), "predict_clean": predict_clean.cpu().numpy(), "noise": noise.cpu().numpy()}
        if predict_noise is not None:
            out["predict_noise"] = predict_noise.cpu().numpy()
        out["noisy"] = noisy.cpu().numpy()
        return out

    def validation_step(self, val_batch, val_idx):
        from .network_functions import val_batch_2_metric_loss
        r = val_batch_2_metric_loss(self, val_batch, val_idx, dtype=self._step_dtype, variant=self.variant, metrics=self.step_metrics)
        if self._two_mask():
            noise_loss, speech_loss, val_loss, pesq_av, stoi_av, predict_noise_audio, predict_clean_audio, noise_audio, noisy_audio, clean_audio = r
            metrics = {'val_loss': val_loss.detach(), 'val_noise_loss': noise_loss.detach(), 'val_speech_loss': speech_loss.detach(),
                       'val_pesq': torch.tensor(pesq_av), 'val_stoi': torch.tensor(stoi_av)}
            output = self._audio_dict(clean_audio, predict_clean_audio, noise_audio, noisy_audio, predict_noise_audio)
            check = val_loss
        else:
            speech_loss, pesq_av, stoi_av, predict_clean_audio, noise_audio, noisy_audio, clean_audio = r
            metrics = {'val_speech_loss': speech_loss.detach(), 'val_pesq': torch.tensor(pesq_av), 'val_stoi': torch.tensor(stoi_av)}
            output = self._audio_dict(clean_audio, predict_clean_audio, noise_audio, noisy_audio)
            check = speech_loss
        if torch.any(torch.isnan(check)):
            print("found a NaN in val loss!")
            return None
        return output, metrics

    def test_step(self, test_batch, test_idx):
        from .network_functions import test_batch_2_metric_loss
        r = test_batch_2_metric_loss(self, test_batch, test_idx, dtype=self._step_dtype, variant=self.variant, metrics=self.step_metrics)
        if self._two_mask():
            noise_loss, speech_loss, test_loss, pesq_av, stoi_av, predict_noise_audio, predict_clean_audio, \
                noise_audio, noisy_audio, clean_audio, id, start_point = r
            metrics = {'test_loss': test_loss, 'test_noise_loss': noise_loss, 'test_speech_loss': speech_loss,
                       'test_pesq': pesq_av, 'test_stoi': stoi_av}
            output = self._audio_dict(clean_audio, predict_clean_audio, noise_audio, noisy_audio, predict_noise_audio)
        else:
            speech_loss, pesq_av, stoi_av, predict_clean_audio, noise_audio, noisy_audio, clean_audio = r
            metrics = {'test_speech_loss': speech_loss, 'test_pesq': pesq_av, 'test_stoi': stoi_av}
            output = self._audio_dict(clean_audio, predict_clean_audio, noise_audio, noisy_audio)
        return output, metrics


    def _epoch_end(self, step_outputs, prefix):
        """validation_epoch_end / test_epoch_end (c_network.py:304-335, 374-398): average the per-batch metrics, log sample audio
        (network_functions.epoch_end), log and return the averages."""
        from .network_functions import epoch_end
        audio = [o[0] for o in step_outputs]
        metrics_list = [o[1] for o in step_outputs]
        mean = lambda k: torch.stack([torch.as_tensor(x[k]) for x in metrics_list]).float().mean()   # noqa: E731
        metrics = {}
        if self._two_mask():
            metrics[f'{prefix}_loss'], metrics[f'{prefix}_noise_loss'] = mean(f'{prefix}_loss'), mean(f'{prefix}_noise_loss')
        for k in ('speech_loss', 'pesq', 'stoi'):
            metrics[f'{prefix}_{k}'] = mean(f'{prefix}_{k}')
        metrics['step'] = self.current_epoch
        if self.logger is not None:
            epoch_end(self, audio, prefix, variant=self.variant)
        self.log_dict(metrics, on_epoch=True)
        return metrics

    def validation_epoch_end(self, validation_step_outputs):
        return self._epoch_end(validation_step_outputs, "val")

    def test_epoch_end(self, test_step_outputs):
        return self._epoch_end(test_step_outputs, "test")

    def on_after_backward(self):
        """c_network.py:401-417: every 25 steps log the mean and the norm of all gradients."""
        step = getattr(getattr(self, "trainer", None), "global_step", self.global_step)
        if step % 25 == 0 and self.logger is not None:
            vals = torch.cat([(p.grad if p.grad is not None else torch.zeros(1, device=p.device)).flatten().float() for p in self.parameters()])
            self.logger.experiment.add_scalar("grad val avg", torch.mean(vals), global_step=step)
            self.logger.experiment.add_scalar("grad norm", torch.linalg.norm(vals), global_step=step)


class ComplexLSTM(torch.nn.Module):
    def __init__(self, input_size, hidden_size, num_layers, bidirectional, batch_first, projection_dim=None):
        super(ComplexLSTM, self).__init__()
        self.input_dim = input_size
        self.rnn_units = hidden_size
        self.real_lstm = torch.nn.LSTM(input_size=self.input_dim, hidden_size=self.rnn_units, num_layers=num_layers,
                                       bidirectional=bidirectional, batch_first=batch_first)
        self.imag_lstm = torch.nn.LSTM(input_size=self.input_dim, hidden_size=self.rnn_units, num_layers=num_layers,
                                       bidirectional=bidirectional, batch_first=batch_first)
        ndir = 2 if bidirectional else 1
        if projection_dim is not None:
            self.projection_dim = projection_dim
            self.r_trans = torch.nn.Linear(self.rnn_units * ndir, self.projection_dim)
            self.i_trans = torch.nn.Linear(self.rnn_units * ndir, self.projection_dim)
        else:
            self.projection_dim = None
        self._cfg = (num_layers, bool(bidirectional), bool(batch_first))
        self._cache = _PackedCache()

    def forward(self, inputs):
        L.require_cuda(inputs)
        if self._cfg != (2, True, True) or self.projection_dim is not None or (self.input_dim, self.rnn_units) != (128, 64):
            raise NotImplementedError("dcsnet_b200.ComplexLSTM: the configuration built at c_network.py:117-123 only "
                                      "(128 -> 64, 2 layers, bidirectional, batch_first, no projection)")
        sd = {"l." + k: v for k, v in self.state_dict().items()}
        key = _version_key(*self.parameters()) + (str(inputs.device),)
        w = self._cache.get(key, lambda: packing.pack_lstm(sd, "l.", inputs.device))
        B, S, D = inputs.shape
        x = torch.view_as_real(inputs.contiguous())
        y = torch.empty(B, S, 2 * self.rnn_units, 2, dtype=torch.float32, device=inputs.device)
        ws = torch.empty(ops.clstm_workspace_bytes(B, S) // 4, dtype=torch.float32, device=inputs.device)
        ops.clstm(x, y, w, ws)
        return torch.view_as_complex(y)

    def flatten_parameters(self):
        self.imag_lstm.flatten_parameters()
        self.real_lstm.flatten_parameters()


class ComplexChannelAttention(torch.nn.Module):
    def __init__(self, no_channels, reduction_ratio):
        super(ComplexChannelAttention, self).__init__()
        self.avg_pool = ComplexAdaptiveAvgPool2d(1)
        self.max_pool = ComplexAdaptiveMaxPool2d(1)
        self.fc = torch.nn.Sequential(
            ComplexConv2d(no_channels, max(no_channels // reduction_ratio, 1), kernel_size=1, bias=False),
            ComplexReLU(),
            ComplexConv2d(max(no_channels // reduction_ratio, 1), no_channels, kernel_size=1, bias=False))
        self.sigmoid = ComplexSigmoid()
        self._cache = _PackedCache()

    def forward(self, x):
        """Returns the (B,C,1,1) complex gate sigmoid_c(fc(avg) + fc("max" = avg))."""
        L.require_cuda(x)
        xr = to_cl(x)
        B, H, W, Cn, _ = xr.shape
        sd = {"a." + k: v for k, v in self.state_dict().items()}
        ca = self._cache.get(_version_key(*self.parameters()) + (str(x.device),),
                             lambda: packing.pack_channel_attention(sd, "a.", x.device))
        sums = ops.zero_(torch.empty(B, Cn, 2, dtype=torch.int64, device=x.device))
        gate = torch.empty(B, Cn, 2, dtype=torch.float32, device=x.device)
        ops.chan_pool(xr, sums)
        ops.chan_gate(sums, H * W, ca, gate)
        return torch.view_as_complex(gate).view(B, Cn, 1, 1)


class ComplexSpatialAttention(torch.nn.Module):
    def __init__(self, kernel_size):
        super(ComplexSpatialAttention, self).__init__()
        self.conv1 = ComplexConv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = ComplexSigmoid()
        self._cache = _PackedCache()

    def forward(self, x):
        """Returns the (B,1,H,W) complex gate sigmoid_c(conv7x7([mean_c x, max_c Re x + j max_c Im x]))."""
        L.require_cuda(x)
        xr = to_cl(x)
        B, H, W, Cn, _ = xr.shape
        sd = {"s." + k: v for k, v in self.state_dict().items()}
        w7 = self._cache.get(_version_key(*self.parameters()) + (str(x.device),),
                             lambda: packing.pack_spatial_attention(sd, "s.", x.device))
        stats = torch.empty(B, H * W, 4, dtype=torch.float32, device=x.device)
        gate = torch.empty(B, H, W, dtype=torch.complex64, device=x.device)
        ops.spat_stats(xr, None, stats)
        ops.spat_apply(xr, None, stats, w7, None, gate_out=gate)
        return gate.view(B, 1, H, W)


class C_NETWORK(_StepMixin, _Base):
    """c_network.py:87-226.  Extra (non-reference) attribute: `compute_mode` in {'fp32', 'fp16', 'bf16'} (engine.MODES) selects the
    CUDA-core fp32 GEMMs (<=1e-5) or the tcgen05 GEMMs on fp16 storage (<=2e-3; 'bf16' = the wide-range variant, ~3.6e-3 on
    the randomised-BN parity state); default 'fp32' = the reference's precision=32."""

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
        self.initial_batchnorm = ComplexBatchNorm2d(max(ch[0] // 2, 1))
        for i in range(n_layers):
            cin = 1 if i == 0 else ch[i] // 2
            self.encoder.append(torch.nn.Sequential(
                ComplexConv2d(in_channels=cin, out_channels=ch[i + 1] // 2, kernel_size=config.kernel_sizeE[i],
                              stride=config.strideE[i], padding=config.paddingE[i]),
                ComplexBatchNorm2d(ch[i + 1] // 2),
                config.CactivationE()))

        self.lstm = ComplexLSTM(input_size=ch[4], hidden_size=ch[4] // 2, num_layers=hp['lstm_layers'],
                                bidirectional=hp['lstm_bidir'], batch_first=True)
        self.fc = ComplexLinear(ch[5] // 2, ch[5] // 2)

        for i in range(n_layers):
            cin_full = ch[n_layers - i]
            cat_channels = (cin_full + cin_full) // 2
            half = cin_full // 2
            cout = max(ch[n_layers - 1 - i] // 2, 1)
            convt = ComplexConvTranspose2d(cat_channels, cout, kernel_size=config.kernel_sizeD[i], stride=config.strideD,
                                           padding=config.paddingD[i])
            if i == n_layers - 1:
                self.decoder.append(convt)
            else:
                self.decoder.append(torch.nn.Sequential(convt, ComplexBatchNorm2d(ch[n_layers - 1 - i] // 2),
                                                        config.CactivationD()))
            self.skip_attention.append(ComplexChannelAttention(half, hp['channel_attention_reduction_ratio']))
            self.skip_attention.append(ComplexSpatialAttention(hp['spatial_attention_kernel_size']))
            self.decoder_attention.append(ComplexChannelAttention(cout, hp['channel_attention_reduction_ratio']))
            self.decoder_attention.append(ComplexSpatialAttention(hp['spatial_attention_kernel_size']))

        self.dropout_conv = torch.nn.Dropout(hp['dropout_conv'])
        self.dropout_fc = torch.nn.Dropout(hp['dropout_fc'])
        self.weights_init()
        self.compute_mode = "fp32"
        self.exact_polar = False
        self._packed = {}
        self._plans = {}

    def weights_init(self):
        init = self.hparams['initialisation_distribution']
        for m in self.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
                init(m.weight)

    # ------------------------------------------------------------------ fused plan management
    def _state_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def packed_for(self, device, mode=None):
        mode = mode or self.compute_mode
        key = (str(device), mode)
        skey = self._state_key()
        hit = self._packed.get(key)
        if hit is None or hit[0] != skey:
            self._packed[key] = (skey, PackedNet(self, device, mode, self.hparams['no_of_layers']))
            self._plans = {k: v for k, v in self._plans.items() if k[:2] != key}
        return self._packed[key][1]

    def plan_for(self, x, variant="dcs"):
        B, F, T = x.shape
        pk = self.packed_for(x.device)
        key = (str(x.device), self.compute_mode, B, F, T, variant, self.exact_polar)
        if key not in self._plans or self._plans[key].pk is not pk:
            self._plans[key] = ForwardPlan(pk, B, T, F, variant=variant, atan2_eps=self.hparams['atan2_eps'],
                                           exact_polar=self.exact_polar)
        return self._plans[key]

    def forward(self, x):
        """x: (B,F,T) complex64 -> bound_cRM(decoder output), squeezed like c_network.py:224-226."""
        if self.training:
            raise NotImplementedError("dcsnet_b200.C_NETWORK.forward has no autograd twin: in train mode use training_step(batch, idx) + "
                                      "train_step.backward() + train_step.optimizer_step() (dcsnet_b200.train_engine.TrainStep: batch-statistic "
                                      "BatchNorm, dropout, the whole backward pass and Adam-amsgrad as kernels), or call .eval()")
        L.require_cuda(x)
        if x.dtype != torch.complex64 or x.dim() != 3:
            raise RuntimeError("C_NETWORK.forward expects a complex64 (B,F,T) spectrogram")
        out = self.plan_for(x.contiguous()).enhance_spec(x.contiguous())["net_out"]
        return torch.squeeze(out.clone())
