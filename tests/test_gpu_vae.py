"""GPU parity of the KL-VAE engine (odise_b200/vae.py) vs oracle/ldm.py::encoder_features /
decoder_features (which equal the reference's LdmExtractor.encoder_forward / decoder_forward, tests/test_oracle_cpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def test_vae_taps(cuda):
    from odise_b200 import spec
    from odise_b200.vae import VAEEngine
    from oracle import ldm
    sd = spec.synth_state_dict(spec.vae_params(), seed=3)
    with torch.device("meta"):
        m = ldm.AutoencoderKL()
    m.load_state_dict({k[len(spec.VAE_PREFIX):]: v for k, v in sd.items()}, assign=True)
    m.eval()
    B, H = 1, 256
    g = torch.Generator().manual_seed(8)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    with torch.no_grad():
        lat, ef = ldm.encoder_features(m, img)
        df = ldm.decoder_features(m, lat)
    eng = VAEEngine(sd, cuda, nmma=3)
    enc = eng.encode(img.permute(0, 2, 3, 1).reshape(-1, 3).contiguous().to(cuda), B, H, H)
    nchw = lambda t: t[0].view(B, t[1], t[2], -1).permute(0, 3, 1, 2).cpu()
    assert _rel(nchw(enc["enc5"]), ef[0]) < 1e-3
    assert _rel(nchw(enc["enc7"]), ef[1]) < 1e-3
    assert _rel(nchw(enc["latent"]), lat) < 1e-3
    # decoder taps from the ORACLE latent (isolates the decoder)
    dec = eng.decode_taps(lat.permute(0, 2, 3, 1).reshape(-1, 4).contiguous().to(cuda), B, lat.shape[2], lat.shape[3])
    torch.cuda.synchronize()
    assert _rel(nchw(dec["dec2"]), df[0]) < 1e-3
    assert _rel(nchw(dec["dec5"]), df[1]) < 1e-3


def test_image_crops(cuda):
    from odise_b200 import ops
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (2, 3, 96, 128), generator=g, dtype=torch.uint8)
    boxes = torch.tensor([[0, 0, 0], [0, 32, 64], [1, 16, 8]], dtype=torch.int32)
    out = ops.image_crops(img.to(cuda), boxes.to(cuda), 3, 96, 128, 64, 64).view(3, 64, 64, 3).cpu()
    for i, (im, y, x) in enumerate(boxes.tolist()):
        want = ((img[im, :, y:y + 64, x:x + 64].float() / 255.0) - 0.5) / 0.5
        assert torch.equal(out[i], want.permute(1, 2, 0))
