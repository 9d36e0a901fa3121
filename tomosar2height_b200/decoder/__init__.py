from . import pixel

# same key as the reference registry (tomosar2height/decoder/__init__.py:4-6)
decoder_dict = {
    'pixel': pixel.PixelwiseDecoder,
}
