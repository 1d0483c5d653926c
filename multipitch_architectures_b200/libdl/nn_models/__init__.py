from .basic_cnns import basic_cnn_segm_sigmoid, deep_cnn_segm_sigmoid
from .unet_cnns import (double_conv, transformer_enc_layer, simple_u_net_largekernels, simple_u_net_doubleselfattn,
                        simple_u_net_doubleselfattn_twolayers, simple_u_net_polyphony_classif_softmax,
                        blstm_temporal_enc_layer, u_net_blstm_varlayers)
