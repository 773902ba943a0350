from neuralmonkey_b200.decoders.decoder import Decoder
