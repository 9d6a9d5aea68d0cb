"""The slice of the reference's yacs config the rollout path reads.

Key names and values follow prosim_demo/cfg/waymo_demo.yaml (the only released
model shape) layered over prosim/config/default.py; ``get_config`` mirrors the
reference's entry point of the same name (config/default.py:690-733) including the
derived ``TARGET.ELEMENTS += ',xd,yd'`` patch for PRED_VEL (:725-730).  A real yacs
node from the reference can be passed to ``ProSimB200`` instead: only attribute
access is used.
"""
import copy


class Config(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split('.')
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = val


def _wrap(d):
    return Config({k: _wrap(v) if isinstance(v, dict) else v for k, v in d.items()})


_ATTN = dict(LEARNABLE_PE=False, NUM_LAYER=6, NUM_HEAD=8, FF_DIM=16, DROPOUT=0.1, PE_NUM_FREQ=64)

_DEFAULTS = {
    'TASK': {'TYPES': ['motion_pred'], 'MOTION_PRED': {'PROMPT': 'agent_status'}},
    'PROMPT': {
        'AGENT_STATUS': {'USE_VEL': True, 'USE_EXTEND': True, 'USE_AGENT_TYPE': True},
        'CONDITION': {'TYPES': [], 'EVAL_COND_SETS': [],
                      'MOTION_TAG': {'USED_TAGS': ['Accelerate', 'Decelerate', 'KeepSpeed', 'Stopping', 'LeftLaneChange',
                                                   'RightLaneChange', 'KeepLane', 'LeftTurn', 'RightTurn', 'Straight',
                                                   'Parked']}},
    },
    'ROLLOUT': {
        'PARALLEL_NUM': 1,
        'POLICY': {'REPLAN_FREQ': 10, 'TOP_K': 1, 'TOP_K_TRAIN': 1, 'MAX_STEPS': 80},
    },
    'LOSS': {'ROLLOUT_TRAJ': {'USE_GOAL_PRED_LOSS': True}},
    'DATASET': {
        'USE_PED_CYCLIST': True,
        'MOTION': {'DT': 0.1},
        'FORMAT': {
            'MAP': {'MAX_POINTS': 2048, 'WITH_TYPE_EMB': True, 'WITH_DIR': True},
            'TARGET': {'SAMPLE_RATE': 10, 'STEPS': 10, 'ELEMENTS': 'x,y,h', 'TAIL_PADDING': True},
            'HISTORY': {'ELEMENTS': 'x,y,s,c,xd,yd,xdd,ydd', 'STEPS': 11, 'WITH_EXTEND': True,
                        'WITH_AGENT_TYPE': True, 'WITH_TIME_EMB': True},
        },
    },
    'MODEL': {
        'TYPE': 'prosim_b200',
        'BPTT': False,
        'HIDDEN_DIM': 128,
        'REL_POS_EDGE_FUNC': 'radius',
        'OBS_UPDATE': {'ATTN_UPDATE': False, 'FUSION': 'replace'},
        'MAP_ENCODER': {'POINTNET': {'NUM_MLP_LAYERS': 5, 'NUM_PRE_LAYERS': 3}},
        'OBS_ENCODER': {'POINTNET': {'NUM_MLP_LAYERS': 3, 'NUM_PRE_LAYERS': 1}},
        'SCENE_ENCODER': {'TYPE': 'attn_fusion_relpe', 'MAP_TYPE': 'pointnet', 'OBS_TYPE': 'pointnet',
                          'ATTN': dict(_ATTN, MAX_NUM_NEIGH=32, AGENT_RADIUS=100, SCENE_RADIUS=50)},
        'DECODER': {'TYPE': 'attn_fusion_relpe', 'GOAL_PRED': {'ENABLE': False, 'K': 1},
                    'ATTN': dict(_ATTN, SCENE_RADIUS=300, PROMPT_RADIUS=300, MAX_NUM_NEIGH=512)},
        'POLICY': {'TYPE': 'rel_pe_temporal',
                   'ACT_DECODER': {
                       'TYPE': 'policy_no_rnn', 'RANDOM_NOISE_STD': 0.0,
                       'MCG': {'LAYER': 3},
                       'TRAJ': {'K': 1, 'PRED_GMM': False, 'PRED_VEL': True, 'PRED_MODE': 'anchor'},
                       'CONTEXT': {'GOAL': False, 'EMD': True, 'GT_GOAL': False, 'USE_POSE_EMB': False},
                       'ATTN': dict(_ATTN, AGENT_RADIUS=100, MAP_RADIUS=50, MAX_NUM_NEIGH=768, NOT_USE_MAP=False)}},
        'CONDITION_TRANSFORMER': {'USE_TEMPORAL_ENCODING': True, 'ATTN_TYPE': 'gnn', 'NLAYER': 3, 'NHEAD': 8,
                                  'FF_DIM': 16, 'DROPOUT': 0.1, 'COND_POOL_FUNC': 'mean',
                                  'CONDITION_LOCATIONS': ['policy_decoder'], 'USE_PLACEHOLDER': True,
                                  'PE': {'ENABLE': True},
                                  'CONDITION_ENCODER': {'DRAG_POINTS': {'NUM_PRE_LAYERS': 1, 'NUM_MLP_LAYERS': 3}}},
    },
}


def get_config(config_paths=None, opts=None, cluster='local'):
    """Defaults of the released demo model; ``opts`` is the reference's flat [key, value, ...] list."""
    cfg = _wrap(copy.deepcopy(_DEFAULTS))
    if config_paths:
        raise NotImplementedError('yaml overlays are out of scope: pass opts or a reference yacs node')
    if opts:
        cfg.merge_from_list(list(opts))
    if cfg.MODEL.POLICY.ACT_DECODER.TRAJ.PRED_VEL:
        if 'xd,yd' not in cfg.DATASET.FORMAT.TARGET.ELEMENTS:
            cfg.DATASET.FORMAT.TARGET.ELEMENTS = cfg.DATASET.FORMAT.TARGET.ELEMENTS + ',xd,yd'
    return cfg


def check_supported(cfg):
    """Raise if a config selects a model variant outside the built path."""
    m = cfg.MODEL
    problems = []
    if m.HIDDEN_DIM != 128:
        problems.append('MODEL.HIDDEN_DIM must be 128')
    if m.REL_POS_EDGE_FUNC != 'radius':
        problems.append("MODEL.REL_POS_EDGE_FUNC must be 'radius'")
    if m.OBS_UPDATE.FUSION not in ('replace', 'mlp'):
        problems.append("MODEL.OBS_UPDATE.FUSION must be 'replace' or 'mlp'")
    for name, a in (('SCENE_ENCODER', m.SCENE_ENCODER.ATTN), ('DECODER', m.DECODER.ATTN),
                    ('POLICY.ACT_DECODER', m.POLICY.ACT_DECODER.ATTN)):
        if a.LEARNABLE_PE or a.NUM_HEAD != 8 or a.FF_DIM != 16:
            problems.append(f'MODEL.{name}.ATTN must be fixed PE, 8 heads x 16')
    t = m.POLICY.ACT_DECODER.TRAJ
    if t.K != 1 or t.PRED_GMM or not t.PRED_VEL or t.PRED_MODE != 'anchor':
        problems.append('MODEL.POLICY.ACT_DECODER.TRAJ must be K=1, anchor, PRED_VEL, no GMM')
    if m.DECODER.GOAL_PRED.ENABLE:
        problems.append('MODEL.DECODER.GOAL_PRED must be disabled')
    if set(cfg.PROMPT.CONDITION.TYPES) - {'goal', 'v_action_tag', 'drag_point'}:
        problems.append("PROMPT.CONDITION.TYPES may only contain 'goal', 'v_action_tag', 'drag_point' (no v2v / text conditions)")
    if len(set(cfg.PROMPT.CONDITION.TYPES)) != len(cfg.PROMPT.CONDITION.TYPES):
        problems.append('PROMPT.CONDITION.TYPES has duplicates')
    if cfg.PROMPT.CONDITION.TYPES and m.CONDITION_TRANSFORMER.COND_POOL_FUNC != 'mean':
        problems.append("MODEL.CONDITION_TRANSFORMER.COND_POOL_FUNC must be 'mean'")
    # the kernels hard-code the released data format: 11 history steps x 24 features per agent (8 motion elements + extent
    # + type one-hot + time one-hot), 10-step targets (x, y, h, xd, yd), dt = 0.1 s, 19 vectors x 11 features per polyline
    f = cfg.DATASET.FORMAT
    if f.HISTORY.ELEMENTS != 'x,y,s,c,xd,yd,xdd,ydd' or f.HISTORY.STEPS != 11 or not (f.HISTORY.WITH_EXTEND and
                                                                                  f.HISTORY.WITH_AGENT_TYPE and f.HISTORY.WITH_TIME_EMB):
        problems.append('DATASET.FORMAT.HISTORY must be the released layout (x,y,s,c,xd,yd,xdd,ydd + extent + type + time, 11 steps)')
    if f.TARGET.STEPS != 10 or f.TARGET.SAMPLE_RATE != 10 or not f.TARGET.ELEMENTS.startswith('x,y,h'):
        problems.append('DATASET.FORMAT.TARGET must be 10 steps of x,y,h(,xd,yd) at SAMPLE_RATE 10')
    if not (f.MAP.WITH_TYPE_EMB and f.MAP.WITH_DIR):
        problems.append('DATASET.FORMAT.MAP must carry the type embedding and direction (11 features per vector)')
    if abs(float(cfg.DATASET.MOTION.DT) - 0.1) > 1e-12:
        problems.append('DATASET.MOTION.DT must be 0.1')
    if cfg.ROLLOUT.POLICY.REPLAN_FREQ != 10:
        problems.append('ROLLOUT.POLICY.REPLAN_FREQ must be 10')
    if problems:
        raise NotImplementedError('; '.join(problems))
