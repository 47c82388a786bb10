"""The Config.py knobs the hot path reads, with the reference's names, meaning and defaults
(/root/reference/ga3c/Config.py; line numbers in the comments).  The reference's `Config` class can
be passed to `Network(..., config=Config)` instead: only attribute access by these names is used.
"""


class Config:
    # input geometry -- Config.py:90-92
    STACKED_FRAMES = 4
    IMAGE_WIDTH = 84
    IMAGE_HEIGHT = 84

    # workers / device -- Config.py:53-62
    AGENTS = 32
    PREDICTORS = 2
    TRAINERS = 2
    DEVICE = 'gpu:0'

    # returns -- Config.py:73-83
    DISCOUNTING = True
    DISCOUNT = 0.99
    REWARD_CLIPPING = True
    USE_INTERMEDIATE_REWARD = False
    REWARD_MIN = -1
    REWARD_MAX = 1

    # queues / batching -- Config.py:86-87, :124
    MAX_QUEUE_SIZE = 100
    PREDICTION_BATCH_SIZE = 128
    TRAINING_MIN_BATCH_SIZE = 0

    # loss -- Config.py:99-100, :122, :153-156
    BETA_START = 0.01
    BETA_END = 0.01
    LOG_EPSILON = 1e-6
    MIN_POLICY = 0.0
    USE_LOG_SOFTMAX = False

    # low-dimensional networks -- Config.py:107 (NetworkVP_discrate builds one dense layer per entry, every one from x)
    DENSE_LAYERS = (10, 10, 10, 10)

    # optimizer -- Config.py:111-120, :194-195
    RMSPROP_DECAY = 0.99
    RMSPROP_MOMENTUM = 0.0
    RMSPROP_EPSILON = 0.1
    DUAL_RMSPROP = False
    USE_GRAD_CLIP = False
    GRAD_CLIP_NORM = 40.0
    LEARNING_RATE_START = 0.0003
    LEARNING_RATE_END = 0.0003

    # rollout length -- Config.py:193 (upstream NVlabs value 5 is commented at :77)
    TIME_MAX = 1000

    # flow control -- Config.py:40-46, :127-139, :185-188
    PLAY_MODE = False
    TRAIN_MODELS = True
    LOAD_CHECKPOINT = False
    LOAD_EPISODE = 0
    SAVE_MODELS = True
    TENSORBOARD = True
    TENSORBOARD_UPDATE_FREQUENCY = 1000
    NETWORK_NAME = 'network'
    USE_REPLAY_MEMORY = False
    USE_NETWORK_TESTER = False
    RANDOM_SEED = 12345

    # annealing -- Config.py:196-197 (episodes over which lr / beta move from START to END)
    EPISODES = 40000
    ANNEALING_EPISODE_COUNT = 40000


def annealed(cfg, episode_count):
    """(learning_rate, beta) as Server.main sets them on the model every 10 ms (Server.py:168-175): linear in the episode
    count, clamped at ANNEALING_EPISODE_COUNT - 1.  `model.learning_rate, model.beta = annealed(Config, episodes)`."""
    lr_mult = (cfg.LEARNING_RATE_END - cfg.LEARNING_RATE_START) / cfg.ANNEALING_EPISODE_COUNT
    beta_mult = (cfg.BETA_END - cfg.BETA_START) / cfg.ANNEALING_EPISODE_COUNT
    step = min(episode_count, cfg.ANNEALING_EPISODE_COUNT - 1)
    return cfg.LEARNING_RATE_START + lr_mult * step, cfg.BETA_START + beta_mult * step
