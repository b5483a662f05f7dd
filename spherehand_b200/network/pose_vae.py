"""Drop-in mirror of the hot-path part of /root/reference/network/pose_vae.py: `PoseVae` with the reference's layer
names (base.0/1/3/4, mu, logvar, decoder.0/1/3/4/6 — `mesh/model/pose_vae.pth` loads unchanged) and `prior_loss`
(:81-89) as ONE fused kernel (csrc/pose_vae.cu).  The training script of that file (:140-189) is out of scope.
"""
import torch
import torch.nn as nn

from .. import ops
from ..mesh.render import _ScalarLossFunction


class PoseVae(nn.Module):
    def __init__(self, pose_fea, latent_fea, model_path=None):
        super().__init__()
        if (pose_fea, latent_fea) != (123, 32):
            raise ValueError('the fused prior kernel is built for the 41x3 -> 32 VAE the reference ships (pose_vae.pth)')
        self.pose_fea = pose_fea
        self.latent_fea = latent_fea
        self.base = nn.Sequential(nn.Linear(pose_fea, 256), nn.GroupNorm(16, 256), nn.ReLU(),
                                  nn.Linear(256, 256), nn.GroupNorm(16, 256), nn.ReLU())
        self.mu = nn.Linear(256, latent_fea)
        self.logvar = nn.Linear(256, latent_fea)
        self.decoder = nn.Sequential(nn.Linear(latent_fea, 256), nn.GroupNorm(16, 256), nn.ReLU(),
                                     nn.Linear(256, 256), nn.GroupNorm(16, 256), nn.ReLU(), nn.Linear(256, pose_fea))
        self._blob = None
        if model_path is not None:
            check_point = torch.load(model_path, map_location='cpu')
            self.load_state_dict(check_point['network_state_dict'])
            for param in self.parameters():
                param.requires_grad = False

    def _weights(self, device):
        """Weights packed for the kernel; re-packed when the module is moved or its state_dict reloaded."""
        key = (str(device), tuple(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._blob is None or self._blob[0] != key:
            self._blob = (key, ops.vae_blob_from_state_dict(self.state_dict(), device))
        return self._blob[1]

    def prior_loss(self, x, eps=None):
        """x [..., 123] (= xyz / 100): MSE(x, recon) + KLD with z = mu + eps * 0.1 * exp(logvar / 2); gradient w.r.t. x.
        eps ~ N(0,1) [M,32] is drawn with torch.randn like the reference (:49-52) unless given."""
        flat = x.reshape(-1, self.pose_fea)
        m = flat.shape[0]
        if eps is None:
            eps = torch.randn(m, self.latent_fea, device=x.device)
        loss3, grad = ops.vae_prior_fwdbwd(flat.detach().contiguous().float(), eps.contiguous().float(), self._weights(x.device))
        return _ScalarLossFunction.apply(x, loss3[0], grad.reshape(x.shape))
