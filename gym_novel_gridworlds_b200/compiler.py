"""Chain compiler: wrapper chain (config builders) -> one flat `ngw_config` table.

For every external action id the compiler asks the outermost wrapper which terminal opcode the id
reaches and which pass-through layers wrap it (`_resolve`), exactly retracing the reference's
outermost-first dispatch (SURVEY §8a "stacking rule").  Item/ids tables, masks, recipes, lidar
layout and the reset program are read off the base env AFTER all constructors ran — the same moment
the reference's `step` would read them."""
import ctypes as C

import numpy as np

from . import opcodes as oc
from .capi import ConfigC
from .core import _Invalid


class CompiledConfig(object):
    def __init__(self):
        self.c = ConfigC()
        self.beam_lut = None           # np.int8 array kept alive for c.beam_lut
        self.obs_dim = 0
        self.n_items = 0
        self.item_names = []           # index = item id
        self.external_ids = []
        self.invalid_reasons = {}      # action id -> why the reference would raise
        self.reset_returns = 'dict'
        self.map_size = 10
        self.inv_obs_names = []
        self.lidar_item_names = []

    def fingerprint(self):
        lut = b'' if self.beam_lut is None else self.beam_lut.tobytes()
        raw = bytearray(C.string_at(C.addressof(self.c), C.sizeof(self.c)))
        off = ConfigC.beam_lut.offset
        raw[off:off + 8] = b'\0' * 8
        return bytes(raw) + lut + bytes([self.map_size])


def _item_id(items_id, name):
    return items_id[name] if name in items_id else oc.NONE


def compile_chain(top):
    base = top.unwrapped
    out = CompiledConfig()
    cfg = out.c
    items_id = base.items_id
    n_items = max(items_id.values()) + 1
    if n_items > oc.MAX_ITEMS or len(base.items) > base.max_items:
        # pogostick_v1_env.py:220
        raise AssertionError("Cannot have more than " + str(base.max_items) + " items")
    out.map_size = int(base.map_size)
    if not (5 <= out.map_size <= oc.MAX_MAP_SIZE):
        raise ValueError("map_size %d outside the supported range [5, %d]" % (out.map_size, oc.MAX_MAP_SIZE))
    out.n_items = n_items
    names = [''] * n_items
    for name, idx in items_id.items():
        names[idx] = name
    out.item_names = names
    cfg.n_items = n_items

    # ---- action table ----
    ext = top._external_action_ids()
    out.external_ids = ext
    n_actions = (max(ext) + 1) if ext else 0
    if n_actions > oc.MAX_ACTIONS:
        raise ValueError("more than %d action ids" % oc.MAX_ACTIONS)
    cfg.n_actions = n_actions
    recipes = []
    for a in range(n_actions):
        e = cfg.actions[a]
        try:
            if a not in ext:
                raise _Invalid("id not in the action table")
            entry = top._resolve(a)
        except _Invalid as why:
            out.invalid_reasons[a] = str(why)
            e.op = oc.OP_INVALID
            continue
        e.op, e.arg, e.variant = entry.op, entry.arg & 0xFF, entry.variant
        if entry.op == oc.OP_CRAFT:
            if entry.recipe not in recipes:
                recipes.append(entry.recipe)
            e.arg = recipes.index(entry.recipe)
        if len(entry.layers) > oc.MAX_LAYERS:
            raise NotImplementedError("more than %d pass-through novelties around one action" % oc.MAX_LAYERS)
        for kind in (oc.LAYER_CRATE, oc.LAYER_FIREWALL):
            if entry.layers.count(kind) > 1:
                raise NotImplementedError("the same pass-through novelty twice in one chain")
        if sum(entry.layers.count(k) for k in (oc.LAYER_FENCE_MEDIUM, oc.LAYER_FENCE_HARD)) > 1:
            raise NotImplementedError("two fencerestriction novelties in one chain")
        for i, layer in enumerate(entry.layers):
            e.layers[i] = layer

    if len(recipes) > oc.MAX_RECIPES:
        raise ValueError("more than %d recipes" % oc.MAX_RECIPES)
    cfg.n_recipes = len(recipes)
    for slot, desc in enumerate(recipes):
        r = cfg.recipes[slot]
        if len(desc['inputs']) > oc.MAX_RECIPE_INPUTS:
            raise ValueError("recipe with more than %d ingredient kinds" % oc.MAX_RECIPE_INPUTS)
        r.n_inputs = len(desc['inputs'])
        for i, (item, qty) in enumerate(desc['inputs']):
            r.in_item[i], r.in_qty[i] = item, qty
        r.out_item, r.out_qty = desc['out_item'], desc['out_qty']
        r.needs_table = 1 if desc['needs_table'] else 0
        r.reward_ok = desc['reward_ok']
        r.cost_missing, r.cost_no_table, r.cost_ok = desc['cost_missing'], desc['cost_no_table'], desc['cost_ok']

    # ---- item classes and the literal names the reference compares against ----
    for name in base.unbreakable_items:
        if name in items_id:
            cfg.unbreakable_mask |= 1 << items_id[name]
    for name in base.entities:
        if name in items_id:
            cfg.entity_mask |= 1 << items_id[name]
    for name in base._BREAK_REWARD_ITEMS:
        if name in items_id:
            cfg.break_reward_mask |= 1 << items_id[name]
    cfg.id_wall = _item_id(items_id, 'wall')
    cfg.id_crafting_table = _item_id(items_id, 'crafting_table')
    cfg.id_tree_log = _item_id(items_id, 'tree_log')
    cfg.id_tree_tap = _item_id(items_id, 'tree_tap')
    cfg.id_rubber = _item_id(items_id, 'rubber')
    cfg.id_wool = _item_id(items_id, 'wool')
    cfg.id_string = _item_id(items_id, 'string')
    cfg.id_goal = _item_id(items_id, base.goal_item_to_craft)
    cfg.id_wooden_axe = _item_id(items_id, 'wooden_axe')
    cfg.id_iron_axe = _item_id(items_id, 'iron_axe')
    cfg.id_fire_wall = _item_id(items_id, 'fire_wall')
    cfg.id_crate = _item_id(items_id, 'crate')
    cfg.id_fence = oc.NONE
    node = top
    while node is not base:
        cls = type(node).__name__
        if cls == 'FenceRestriction':
            cfg.id_fence = _item_id(items_id, node.env2.fence_name)
        if cls == 'Crate':
            for name in node.crate_ingredients:
                cfg.crate_add[items_id[str(name)]] += 1
        node = node.env
    cfg.reward_intermediate = int(top.reward_intermediate)
    cfg.reward_done = int(top.reward_done)
    cfg.reward_firewall = -int(top.reward_done) // 2          # novelty_wrappers.py:1187

    # ---- lidar ----
    lidar = top._lidar()
    for i in range(oc.MAX_ITEMS):
        cfg.lidar_slot[i] = -1
    if lidar is not None:
        cfg.n_beams = lidar.num_beams
        cfg.max_range = lidar.max_beam_range
        cfg.n_lidar_items = len(lidar.lidar_items_id)
        for name, slot in lidar.lidar_items_id.items():
            if name in items_id:
                cfg.lidar_slot[items_id[name]] = slot - 1
        out.lidar_item_names = sorted(lidar.lidar_items_id, key=lidar.lidar_items_id.get)
        # inventory tail: sorted names of the LIVE inventory minus the LIVE unbreakable set
        # (observation_wrappers.py:77-78); after a reset the inventory keys are exactly base.items
        tail = [name for name in sorted(base.items) if name not in base.unbreakable_items]
        out.inv_obs_names = tail
        cfg.n_inv_obs = len(tail)
        for i, name in enumerate(tail):
            cfg.inv_obs_item[i] = items_id[name]
        out.beam_lut = np.ascontiguousarray(lidar.beam_lut())
        cfg.beam_lut = out.beam_lut.ctypes.data
        out.obs_dim = cfg.n_lidar_items * cfg.n_beams + cfg.n_inv_obs

    # ---- reset program ----
    prog = top._reset_program()
    if len(prog.place) > oc.MAX_PLACE:
        raise ValueError("more than %d entries in items_quantity" % oc.MAX_PLACE)
    cfg.n_place = len(prog.place)
    for i, (item, qty) in enumerate(prog.place):
        cfg.place_item[i], cfg.place_qty[i] = item, qty
    if len(prog.ops) > oc.MAX_RESET_OPS:
        raise ValueError("more than %d reset post-ops" % oc.MAX_RESET_OPS)
    cfg.n_reset_ops = len(prog.ops)
    replaced_wall = False
    for kind, a, b, lo, hi in prog.ops:
        if kind == oc.RESET_REPLACE and a == cfg.id_wall:
            replaced_wall = True
        if kind == oc.RESET_FENCE and replaced_wall:
            # the reference's Fence.reset then fences around BORDER cells and add_fence_around indexes outside the
            # grid (IndexError, or a silent wrap to the opposite border) — pogostick_v1_env.py:533
            raise NotImplementedError("fence / fencerestriction outside a wall-replacing novelty: the reference "
                                      "raises IndexError in add_fence_around for this chain")
    for i, (kind, a, b, lo, hi) in enumerate(prog.ops):
        op = cfg.reset_ops[i]
        op.kind, op.a, op.b, op.lo, op.hi = kind, a, b, lo, hi
    cfg.reset_obs_after_ops = len(prog.ops) if prog.obs_after is None else prog.obs_after
    out.reset_returns = prog.returns
    return out
