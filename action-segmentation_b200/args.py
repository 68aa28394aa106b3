"""Argument namespace with the defaults of the reference's argparse flags that the HSMM path reads
(/root/reference/src/models/semimarkov/semimarkov_modules.py:53-65, models/semimarkov/semimarkov.py:16-31,
models/model.py:7-24, main.py:60-102).  The reference builds it with argparse in main.py; callers that
drive `SemiMarkovModule` / `SemiMarkovModel` directly (tests, bench.py, smoke) use this instead."""


class HsmmArgs:
    def __init__(self, **kw):
        # SemiMarkovModule.add_args
        self.sm_max_span_length = 20
        self.sm_supervised_state_smoothing = 1e-2
        self.sm_supervised_length_smoothing = 1e-1
        self.sm_supervised_method = "closed-form"
        self.sm_feature_projection = False
        self.sm_init_non_projection_parameters_from = None
        # SemiMarkovModel.add_args
        self.sm_component_model = False
        self.sm_constrain_transitions = False
        self.sm_constrain_with_narration = []
        self.sm_constrain_narration_weight = -1e4
        self.sm_train_discriminatively = False
        self.sm_hidden_markov = False
        self.sm_predict_single = False
        # extensions of this implementation (SemiMarkovModel.add_args)
        self.sm_unsupervised_method = "gradient"
        self.sm_no_device_cache = False
        # models/model.py add_training_args
        self.epochs = 60
        self.batch_accumulation = 1
        self.lr = 5e-3
        self.workers = 0
        self.max_grad_norm = 10
        self.print_every = 100
        self.no_reduce_plateau = False
        self.reduce_plateau_factor = 0.2
        self.reduce_plateau_patience = 1
        self.reduce_plateau_min_lr = 1e-4
        self.train_limit = None
        self.dev_decode_frequency = 1
        # main.py
        self.batch_size = 5
        self.cuda = True
        self.training = "supervised"
        self.annotate_background_with_previous = False
        self.no_merge_classes = False
        self.seed = 0
        self.__dict__.update(kw)
