"""Make the unmodified reference network use the B200 kernels: rebind the module globals the reference resolves at
construction time: `models.ADNMUNet.Mamba2` (looked up inside `create_block`, models/ADNMUNet.py:277),
`models.model_untils.WTConv2d` (models/model_untils.py:17,101) and - for the fused Block - the names
`Block` / `RMSNorm` that `create_block` resolves (models/ADNMUNet.py:278-291) and - for the conv stages -
the names `WTLayer` / `PatchEmbed` / `OutProj` that `Encoder` / `Decoder` / `Refiner` resolve (models/ADNMUNet.py:369-397,
542-567,702-709; star-imported from models.model_untils at :33).  No reference file is edited; call this before
`create_ADNMUNet(...)`."""
import importlib


def install_into_reference(adnmunet_module="models.ADNMUNet", untils_module="models.model_untils",
                           mixer=True, wtconv=True, block=True, stages=True):
    from adnm_unet_b200.mixer import Mamba2
    from adnm_unet_b200.wtconv import WTConv2d
    done = []
    if mixer:
        m = importlib.import_module(adnmunet_module)
        m.Mamba2 = Mamba2
        done.append(adnmunet_module + ".Mamba2")
    if wtconv:
        m = importlib.import_module(untils_module)
        m.WTConv2d = WTConv2d
        done.append(untils_module + ".WTConv2d")
    if block:
        from adnm_unet_b200.block import Block
        from adnm_unet_b200.rmsnorm import RMSNorm
        m = importlib.import_module(adnmunet_module)
        from adnm_unet_b200.attention import StandardAttention
        m.Block, m.RMSNorm, m.StandardAttention = Block, RMSNorm, StandardAttention
        from adnm_unet_b200.block import FeedForward
        importlib.import_module(untils_module).FeedForward = FeedForward      # the FeedForwards of the EncoderToDecoder bridges
        done += [adnmunet_module + ".Block", adnmunet_module + ".RMSNorm", adnmunet_module + ".StandardAttention",
                 untils_module + ".FeedForward"]
    if stages and wtconv:
        from adnm_unet_b200 import convstage
        m = importlib.import_module(adnmunet_module)
        for n in ("WTLayer", "PatchEmbed", "OutProj"):
            setattr(m, n, getattr(convstage, n))
            done.append(adnmunet_module + "." + n)
        # the grouped convs of the EncoderToDecoder bridges: a subclass of the reference's own Conv2dLayer (models/model_untils.py:71-93),
        # resolved by EncoderToDecoder.__init__ from models.model_untils' globals (:623-673)
        u = importlib.import_module(untils_module)
        if not getattr(u.Conv2dLayer, "_adnb200_bridge", False):
            u.Conv2dLayer = convstage.make_bridge_conv_layer(u.Conv2dLayer)
        done.append(untils_module + ".Conv2dLayer")
    return done
