"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional CPU fp32 restatement of ``SmirkGenerator.forward`` (src/smirk_generator.py:51-86) in eval
mode, driven by a reference-format ``state_dict`` (key names per ``_block`` :88-119, ``ResnetBlock``
:121-178).  Tier A: ``oracle/make_golden.py`` checks it against the reference module itself.
"""
import torch
import torch.nn.functional as F


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=eps)


def _block(x, sd, mod, name):
    """conv3x3(p1, no bias) -> BN -> ReLU, twice (smirk_generator.py:88-119)."""
    for i in (1, 2):
        x = F.conv2d(x, sd["%s.%sconv%d.weight" % (mod, name, i)], padding=1)
        x = F.relu(_bn(x, sd, "%s.%snorm%d" % (mod, name, i)))
    return x


def _resblock(x, sd, p):
    """x + [reflpad, conv, BN, ReLU, reflpad, conv, BN](x) (smirk_generator.py:147-178)."""
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[p + ".conv_block.1.weight"])
    y = F.relu(_bn(y, sd, p + ".conv_block.2"))
    y = F.conv2d(F.pad(y, (1, 1, 1, 1), mode="reflect"), sd[p + ".conv_block.5.weight"])
    return x + _bn(y, sd, p + ".conv_block.6")


def generator_forward_ref(sd, x, res_blocks=5):
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    e1 = _block(x, sd, "encoder1", "enc1")
    e2 = _block(F.max_pool2d(e1, 2, 2), sd, "encoder2", "enc2")
    e3 = _block(F.max_pool2d(e2, 2, 2), sd, "encoder3", "enc3")
    e4 = _block(F.max_pool2d(e3, 2, 2), sd, "encoder4", "enc4")
    b = _block(F.max_pool2d(e4, 2, 2), sd, "bottleneck", "bottleneck")
    for i in range(res_blocks):
        b = _resblock(b, sd, "resnet_blocks.%d" % i)
    d = b
    for lvl, skip in ((4, e4), (3, e3), (2, e2), (1, e1)):
        d = F.conv_transpose2d(d, sd["upconv%d.weight" % lvl], sd["upconv%d.bias" % lvl], stride=2)
        d = _block(torch.cat((d, skip), 1), sd, "decoder%d" % lvl, "dec%d" % lvl)
    return torch.sigmoid(F.conv2d(d, sd["conv.weight"], sd["conv.bias"]))
