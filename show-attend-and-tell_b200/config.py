"""Hyper-parameter bag with the reference's field names (config.py:4-44, main.py:35,42).

Only the fields the decode path reads are kept; paths, dataset sizes and CNN options of
the reference Config are out of scope (features are precomputed).  Two fields replace
things the reference derives elsewhere: ``num_ctx``/``dim_ctx`` (model.py:54-59: 196x512
for vgg16; model.py:103-108: 49x2048 for resnet50) and ``eos_id`` (the id of '.' in the
vocabulary, base_model.py:229; 2 in the shipped data/vocabulary.csv).
"""


class Config(object):
    def __init__(self, **overrides):
        # about the model architecture (config.py:8-17)
        self.cnn = 'vgg16'
        self.max_caption_length = 20
        self.dim_embedding = 512
        self.num_lstm_units = 512
        self.num_initalize_layers = 2    # 1 or 2 (spelling as in the reference)
        self.dim_initalize_layer = 512
        self.num_attend_layers = 2       # 1 or 2
        self.dim_attend_layer = 512
        self.num_decode_layers = 2       # 1 or 2
        self.dim_decode_layer = 1024
        # about the weight initialization and regularization (config.py:20-27)
        self.fc_kernel_initializer_scale = 0.08
        self.fc_kernel_regularizer_scale = 1e-4
        self.fc_drop_rate = 0.5
        self.lstm_drop_rate = 0.3
        self.attention_loss_factor = 0.01
        # about the optimization (config.py:30-43)
        self.batch_size = 20
        self.optimizer = 'Adam'          # 'Adam', 'RMSProp', 'Momentum' or 'SGD' (model.py:479-503)
        self.initial_learning_rate = 0.0001
        self.learning_rate_decay_factor = 1.0
        self.num_steps_per_decay = 100000
        self.clip_gradients = 5.0
        self.momentum = 0.0
        self.use_nesterov = True
        self.decay = 0.9
        self.centered = True
        self.beta1 = 0.9
        self.beta2 = 0.999
        self.epsilon = 1e-6
        # The reference hands an Optimizer INSTANCE to tf.contrib.layers.optimize_loss, so its staircase decay
        # (model.py:466-476) only reaches the "learning_rate" summary and the optimizer keeps initial_learning_rate.
        # False reproduces that; True feeds the decayed rate to the optimizer.
        self.apply_learning_rate_decay = False
        # about the saver (config.py:46-49, base_model.py:242-255)
        self.save_period = 1000
        self.save_dir = './models/'
        self.dropout_seed = 0x5A17B200
        # about the vocabulary (config.py:67) and beam search (main.py:35)
        self.vocabulary_size = 5000
        self.beam_size = 3
        self.phase = 'eval'
        # derived by the reference from `cnn`
        self.num_ctx = 196
        self.dim_ctx = 512
        self.eos_id = 2
        for k, v in overrides.items():
            if not hasattr(self, k):
                raise AttributeError("unknown config field %r" % k)
            setattr(self, k, v)
        if 'cnn' in overrides and 'num_ctx' not in overrides and self.cnn == 'resnet50':
            self.num_ctx, self.dim_ctx = 49, 2048
