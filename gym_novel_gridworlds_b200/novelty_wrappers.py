"""Novelties as per-config parameter tables (reference: novelty_wrappers.py:9-1674).

In the reference every novelty is a ``gym.Wrapper`` whose ``step`` re-implements one branch of the
base env in Python.  Here a novelty wrapper does two things only:

1. its constructor applies the SAME table mutations the reference's constructor applies to the base
   env (new item ids, new action ids, recipes, entity / unbreakable sets — these are API-visible
   state that user scripts read back), and
2. it answers three flattening questions for the compiler — which terminal opcode an action id
   reaches through it (``_resolve``), what it appends to the reset program (``_reset_program``), and
   nothing else.

The stacking rule that makes this a table: a *terminal* interceptor handles its action
entirely on the base env, so the OUTERMOST terminal interceptor of an action wins; *pass-through*
interceptors (Crate, FenceRestriction, FireWall) become layers recorded around the terminal opcode.
"""
import numpy as np

from . import spaces
from . import opcodes as oc
from .core import Wrapper, ActionEntry, _Invalid


class _Novelty(Wrapper):
    """Shared helpers.  `self._ids(...)` is the reference's per-step preamble
    (e.g. novelty_wrappers.py:39-43): limited ids if a LimitActions sits anywhere below."""

    def _ids(self, required=(), why=''):
        if hasattr(self, 'limited_actions_id'):
            for name in required:
                assert name in self.limited_actions_id, \
                    "Cannot use " + why + " because you do not have " + name + " in LimitActions"
            return self.limited_actions_id
        return self.actions_id

    @staticmethod
    def _id_of(ids, name):
        if name not in ids:
            raise _Invalid("KeyError: %r" % name)
        return ids[name]

    # ---- constructor-time table edits, phrased once ----
    def _register_item(self, name):
        self.env.items.add(name)
        self.env.items_id.setdefault(name, len(self.items_id))

    def _register_select(self, name):
        self.env.select_actions_id.update({'Select_' + name: len(self.env.actions_id)})
        self.env.actions_id.update(self.env.select_actions_id)

    def _claim_top(self):
        self.unwrapped._top = self


def _axe_break_entry(wrapper, variant):
    return ActionEntry(oc.OP_BREAK, arg=wrapper.items_id[wrapper.axe_name], variant=variant)


def _axe_craft_entry(wrapper):
    """AxeHard.craft / AxetoBreakHard.craft called with the axe only (novelty_wrappers.py:371-436,779-844):
    no 'missing' cost, 600 without a table, 6000 on success, reward_intermediate even on Bow-v1 (Q10)."""
    base = wrapper.unwrapped
    desc = base._recipe_descriptor(wrapper.axe_name, 0, 600.0, 6000.0, wrapper.reward_intermediate)
    # the tree_tap / pogo_stick cost branches of that craft() are unreachable for the axe name
    return ActionEntry(oc.OP_CRAFT, recipe=desc)


class _AxeBase(_Novelty):
    _variant_plain = oc.BRK_AXE
    _why = "breakincrease novelty_arg2"

    def _break_variant(self):
        if getattr(self, 'breakincrease', 'false') == 'true':
            return oc.BRK_AXE_INC
        return self._variant_plain

    def _required(self):
        return ('Break',)

    def _resolve(self, action_id):
        ids = self._ids(self._required(), self._why)
        craft_name = 'Craft_' + self.axe_name
        if self._has_craft and action_id == self._id_of(ids, craft_name):
            return _axe_craft_entry(self)
        if action_id == self._id_of(ids, 'Break'):
            return _axe_break_entry(self, self._break_variant())
        return self.env._resolve(action_id)

    _has_craft = False


class AxeEasy(_AxeBase):
    """Axe starts in the inventory (novelty_wrappers.py:9-35)."""

    def __init__(self, env, axe_material, breakincrease='false'):
        super().__init__(env)
        self.axe_name = axe_material + '_axe'
        self._register_item(self.axe_name)
        self.env.inventory_items_quantity.update({self.axe_name: 1})
        self.env.entities.add(self.axe_name)
        self._register_select(self.axe_name)
        self.breakincrease = breakincrease

    def _reset_program(self):
        prog = self.env._reset_program()
        prog.ops.append((oc.RESET_INVSET, self.items_id[self.axe_name], 0, 1, 0))   # novelty_wrappers.py:33
        return prog


class AxeMedium(_AxeBase):
    """One axe is placed in the map and picked up as an entity (novelty_wrappers.py:117-134)."""

    def __init__(self, env, axe_material, breakincrease='false'):
        super().__init__(env)
        self.axe_name = axe_material + '_axe'
        self.env.add_new_items({self.axe_name: 1})
        self.env.entities.add(self.axe_name)
        self._register_select(self.axe_name)
        self.breakincrease = breakincrease


def _axe_recipe(material):
    if material == 'wooden':
        return {'stick': 2, 'plank': 3}
    if material == 'iron':
        return {'stick': 2, 'iron': 3}
    raise UnboundLocalError("axe_recipe")     # the reference leaves it unbound for other materials


class AxeHard(_AxeBase):
    """Axe must be crafted; its ingredients are placed in the map (novelty_wrappers.py:216-258)."""
    _has_craft = True
    _why = "AxeHard novelty"

    def __init__(self, env, axe_material, breakincrease='false'):
        super().__init__(env)
        self.axe_material = axe_material
        self.axe_name = axe_material + '_axe'
        self._register_item(self.axe_name)
        self.env.inventory_items_quantity.update({self.axe_name: 0})
        self.env.entities.add(self.axe_name)

        recipe = _axe_recipe(axe_material)
        for item, need in recipe.items():
            if item not in self.env.items:
                self.env.add_new_items({item: need})
            else:
                self.env.items_quantity.update({item: self.items_quantity.get(item, 0) + need})
        self.env.recipes.update({self.axe_name: {'input': recipe, 'output': {self.axe_name: 1}}})

        craft = 'Craft_' + self.axe_name
        self.env.craft_actions_id.update({craft: len(self.env.actions_id)})
        self.env.actions_id.update({craft: len(self.env.actions_id)})
        self._register_select(self.axe_name)
        self.env.action_space = spaces.Discrete(len(self.env.actions_id))
        self.breakincrease = breakincrease

    def _required(self):
        return ('Craft_' + self.axe_name, 'Break')


class _AxetoBreakBase(_AxeBase):
    _variant_plain = oc.BRK_AXETOBREAK
    _why = "axetobreak novelty"

    def _break_variant(self):
        return oc.BRK_AXETOBREAK


class AxetoBreakEasy(_AxetoBreakBase):
    """novelty_wrappers.py:439-462."""

    def __init__(self, env, axe_material):
        super().__init__(env)
        self.axe_name = axe_material + '_axe'
        self._register_item(self.axe_name)
        self.env.inventory_items_quantity.update({self.axe_name: 1})
        self.env.entities.add(self.axe_name)
        self._register_select(self.axe_name)

    def _reset_program(self):
        prog = self.env._reset_program()
        prog.ops.append((oc.RESET_INVSET, self.items_id[self.axe_name], 0, 1, 0))   # novelty_wrappers.py:460
        return prog


class AxetoBreakMedium(_AxetoBreakBase):
    """novelty_wrappers.py:537-552."""

    def __init__(self, env, axe_material):
        super().__init__(env)
        self.axe_name = axe_material + '_axe'
        self.env.add_new_items({self.axe_name: 1})
        self.env.entities.add(self.axe_name)
        self._register_select(self.axe_name)


class AxetoBreakHard(_AxetoBreakBase):
    """Agent starts every episode holding the axe ingredients (novelty_wrappers.py:627-673)."""
    _has_craft = True
    _why = "AxetoBreakHard novelty"

    def __init__(self, env, axe_material):
        super().__init__(env)
        self.axe_material = axe_material
        self.axe_name = axe_material + '_axe'
        self._register_item(self.axe_name)
        self.env.inventory_items_quantity.update({self.axe_name: 0})
        self.env.entities.add(self.axe_name)

        recipe = _axe_recipe(axe_material)
        for item in recipe:
            if item not in self.env.items:
                self._register_item(item)
        self.env.inventory_items_quantity.update(recipe)
        self.env.recipes.update({self.axe_name: {'input': recipe, 'output': {self.axe_name: 1}}})
        self._recipe = recipe

        # Craft_<axe> joins actions_id but NOT craft_actions_id (novelty_wrappers.py:658-659)
        self.env.actions_id.update({'Craft_' + self.axe_name: len(self.env.actions_id)})
        self._register_select(self.axe_name)
        self.env.action_space = spaces.Discrete(len(self.env.actions_id))

    def _required(self):
        return ('Craft_' + self.axe_name, 'Break')

    def _reset_program(self):
        prog = self.env._reset_program()
        prog.ops.append((oc.RESET_INVSET, self.items_id[self.axe_name], 0, 0, 0))   # novelty_wrappers.py:668-671
        for item, qty in self._recipe.items():
            prog.ops.append((oc.RESET_INVSET, self.items_id[item], 0, qty, 0))
        return prog


_FENCE_PERCENT = {'easy': (20, 50), 'medium': (50, 90)}


class Fence(_Novelty):
    """Fence blocks around a random share of the placed items after each reset (novelty_wrappers.py:847-889)."""

    def __init__(self, env, difficulty, fence_material):
        super().__init__(env)
        self.fence_name = fence_material + '_fence'
        self._register_item(self.fence_name)
        self._register_select(self.fence_name)
        self.fence_percent_range = _FENCE_PERCENT.get(difficulty, (90, 100))

    def _reset_program(self):
        prog = self.env._reset_program()
        lo, hi = self.fence_percent_range
        prog.ops.append((oc.RESET_FENCE, self.items_id[self.fence_name], self.items_id['wall'], lo, hi))
        prog.returns = 'dict'                  # novelty_wrappers.py:886 returns get_observation()
        return prog


class FenceRestriction(_Novelty):
    """Break is refused near fences until they are broken (novelty_wrappers.py:892-988)."""

    def __init__(self, env, difficulty, fence_material):
        super().__init__(env)
        self.difficulty = difficulty
        self.env2 = Fence(env, 'medium', fence_material=fence_material)
        self._claim_top()

    def _reset_program(self):
        return self.env2._reset_program()

    def _resolve(self, action_id):
        ids = self._ids(('Break',), "fencerestriction novelty")
        entry = self.env._resolve(action_id)
        if action_id == self._id_of(ids, 'Break') and self.difficulty != 'easy':
            entry.layers.insert(0, oc.LAYER_FENCE_MEDIUM if self.difficulty == 'medium' else oc.LAYER_FENCE_HARD)
        return entry


_ADDITEM_PERCENT = {'easy': (1, 10), 'medium': (10, 20)}


class AddItem(_Novelty):
    """A new breakable item fills a random share of the air cells (novelty_wrappers.py:991-1034)."""

    def __init__(self, env, difficulty, item_to_add):
        super().__init__(env)
        self.item_to_add = item_to_add
        self._register_item(item_to_add)
        self._register_select(item_to_add)
        self.item_percent_range = _ADDITEM_PERCENT.get(difficulty, (20, 30))

    def _reset_program(self):
        prog = self.env._reset_program()
        lo, hi = self.item_percent_range
        prog.ops.append((oc.RESET_ADDITEM, self.items_id[self.item_to_add], 0, lo, hi))
        prog.returns = 'dict'                  # novelty_wrappers.py:1031
        return prog


_CRATE_PERCENT = {'easy': (99, 100), 'medium': (50, 90)}


class Crate(_Novelty):
    """Crates (AddItem easy) that drop part of the goal recipe when broken (novelty_wrappers.py:1037-1092).

    `crate_ingredients` is drawn from the global legacy np.random stream with the reference's own
    call sequence (randint, then choice until filled), so a seeded script gets the same crate."""

    def __init__(self, env, difficulty):
        super().__init__(env)
        self.env2 = AddItem(env, 'easy', item_to_add='crate')
        self._claim_top()

        lo, hi = _CRATE_PERCENT.get(difficulty, (10, 50))
        item_percent = np.random.randint(low=lo, high=hi, size=1)[0]
        goal_inputs = self.recipes[self.goal_item_to_craft]['input']
        ingredients = list(goal_inputs)
        wanted = int(np.ceil((item_percent / 100) * sum(goal_inputs.values())))
        self.crate_ingredients = []
        while wanted:
            item = np.random.choice(ingredients, size=1)[0]
            if self.crate_ingredients.count(item) < goal_inputs[item]:
                self.crate_ingredients.append(item)
                wanted -= 1

    def _reset_program(self):
        return self.env2._reset_program()

    def _resolve(self, action_id):
        ids = self._ids(('Break',), "crate novelty")
        entry = self.env._resolve(action_id)
        if action_id == self._id_of(ids, 'Break'):
            entry.layers.insert(0, oc.LAYER_CRATE)
        return entry


_REPLACE_PERCENT = {'easy': (5, 20), 'medium': (40, 90)}


class ReplaceItem(_Novelty):
    """Replace a random share of one item with a new one (novelty_wrappers.py:1095-1148)."""

    def __init__(self, env, difficulty, item_to_replace='wall', item_to_replace_with='brick'):
        super().__init__(env)
        self.item_to_replace = item_to_replace
        self.item_to_replace_with = item_to_replace_with
        assert item_to_replace in self.env.items_id, \
            "Item to replace (" + item_to_replace + ") is not in the original map"
        assert item_to_replace_with not in self.env.items_id, \
            "Item to replace with (" + item_to_replace_with + ") should be a new item"
        self._register_item(item_to_replace_with)
        self._register_select(item_to_replace_with)
        if item_to_replace == 'wall':
            self.env.unbreakable_items.add(item_to_replace_with)
        self.item_percent_range = _REPLACE_PERCENT.get(difficulty, (99, 100))

    def _reset_program(self):
        prog = self.env._reset_program()
        lo, hi = self.item_percent_range
        prog.ops.append((oc.RESET_REPLACE, self.items_id[self.item_to_replace],
                         self.items_id[self.item_to_replace_with], lo, hi))
        prog.returns = 'dict'                  # novelty_wrappers.py:1145
        return prog


class FireWall(_Novelty):
    """Walls become fire_wall; standing next to one ends the episode with -reward_done // 2
    (novelty_wrappers.py:1151-1200)."""

    def __init__(self, env, difficulty='hard'):
        super().__init__(env)
        self.env2 = ReplaceItem(env, difficulty, item_to_replace='wall', item_to_replace_with='fire_wall')
        self._claim_top()

    def _reset_program(self):
        return self.env2._reset_program()

    def _resolve(self, action_id):
        entry = self.env._resolve(action_id)
        entry.layers.insert(0, oc.LAYER_FIREWALL)
        return entry


def remap_action_difficulty(env, difficulty='hard'):
    """Random action-id permutation (novelty_wrappers.py:1203-1227).  Attribute writes on a wrapper
    shadow instead of reaching the base env — kept on purpose."""
    if hasattr(env, 'limited_actions_id'):
        env.set_limited_actions_id(env.remap_action(env.limited_actions_id, 0))
        return env
    if difficulty in ('easy', 'medium'):
        env.manipulation_actions_id = env.remap_action(env.manipulation_actions_id, 0)
        if difficulty == 'medium':
            env.craft_actions_id = env.remap_action(env.craft_actions_id, len(env.manipulation_actions_id))
        env.actions_id.update(env.manipulation_actions_id)
        if difficulty == 'medium':
            env.actions_id.update(env.craft_actions_id)
    else:
        env.actions_id = env.remap_action(env.actions_id, 0)
        env.craft_actions_id = {a: env.actions_id[a] for a in env.actions_id if a.startswith('Craft')}
        env.select_actions_id = {a: env.actions_id[a] for a in env.actions_id if a.startswith('Select')}
    return env


class _NewManipulation(_Novelty):
    _name = None
    _op = None

    def __init__(self, env):
        super().__init__(env)
        self.env.manipulation_actions_id[self._name] = len(self.actions_id)
        self.env.actions_id.update(self.manipulation_actions_id)
        self.action_space = spaces.Discrete(len(self.actions_id))

    def _resolve(self, action_id):
        ids = self._ids((self._name,), "add" + self._name.lower() + " novelty")
        if action_id == self._id_of(ids, self._name):
            return ActionEntry(self._op)
        return self.env._resolve(action_id)


class AddChopAction(_NewManipulation):
    """Chop = Break that yields 2 items at 1.2x the cost (novelty_wrappers.py:1267-1337)."""
    _name, _op = 'Chop', oc.OP_CHOP


class AddJumpAction(_NewManipulation):
    """Jump = move 2 cells ahead, ignoring the cell in between (novelty_wrappers.py:1340-1412)."""
    _name, _op = 'Jump', oc.OP_JUMP


class BreakIncrease(_Novelty):
    """Break yields 2 of `itemtobreakmore` (or of everything) and always rewards (novelty_wrappers.py:1415-1488)."""

    def __init__(self, env, itemtobreakmore=''):
        super().__init__(env)
        self.itemtobreakmore = itemtobreakmore

    def _resolve(self, action_id):
        ids = self._ids(('Break',), "breakincrease novelty")
        if action_id == self._id_of(ids, 'Break'):
            if self.itemtobreakmore == '':
                which = oc.NONE
            else:
                # a name that never equals a front block behaves as "increase nothing"
                which = self.items_id.get(self.itemtobreakmore, 0xFE)
            return ActionEntry(oc.OP_BREAK, arg=which, variant=oc.BRK_INCREASE)
        return self.env._resolve(action_id)


class ExtractIncDec(_Novelty):
    """Extract_* yields double / half (novelty_wrappers.py:1491-1581)."""

    def __init__(self, env, incdec='decrease'):
        super().__init__(env)
        self.incdec = incdec

    def _resolve(self, action_id):
        if hasattr(self, 'limited_actions_id'):
            assert any(a.startswith('Extract') for a in self.limited_actions_id), \
                "Cannot use extractincdec novelty because you do not have Extract action in LimitActions"
            ids = self.limited_actions_id
        else:
            ids = self.actions_id
        if action_id not in ids.values():                      # novelty_wrappers.py:1515 raises ValueError
            raise _Invalid("ValueError: %r is not in list" % (action_id,))
        name = next(k for k, v in ids.items() if v == action_id)
        if not name.startswith('Extract'):
            return self.env._resolve(action_id)
        if self.env_id.startswith('NovelGridworld-Bow'):
            return ActionEntry(oc.OP_EXTRACT_STRING, arg=4 * 2 if self.incdec == 'increase' else 4 // 2)
        if self.env_id.startswith('NovelGridworld-Pogostick'):
            return ActionEntry(oc.OP_EXTRACT_RUBBER, arg=2 if self.incdec == 'increase' else 0)
        raise NotImplementedError(self.env_id)


_NOVELTY_NAMES = ['addchop', 'additem', 'addjump', 'axe', 'axetobreak', 'breakincrease', 'crate', 'extractincdec',
                  'fence', 'fencerestriction', 'firewall', 'remapaction', 'replaceitem']
_WITH_DIFFICULTY = ['additem', 'axe', 'axetobreak', 'crate', 'fence', 'fencerestriction', 'firewall', 'remapaction',
                    'replaceitem']
_AXE = {'easy': AxeEasy, 'medium': AxeMedium, 'hard': AxeHard}
_AXETOBREAK = {'easy': AxetoBreakEasy, 'medium': AxetoBreakMedium, 'hard': AxetoBreakHard}


def inject_novelty(env, novelty_name, difficulty='hard', novelty_arg1='', novelty_arg2=''):
    """Same signature, validation and error types as novelty_wrappers.py:1586-1674."""
    assert novelty_name in _NOVELTY_NAMES, "novelty_name must be one of " + str(_NOVELTY_NAMES)
    if novelty_name in _WITH_DIFFICULTY:
        assert difficulty in ['easy', 'medium', 'hard'], "difficulty must be one of 'easy', 'medium', 'hard'"

    if novelty_name == 'addchop':
        return AddChopAction(env)
    if novelty_name == 'addjump':
        return AddJumpAction(env)
    if novelty_name == 'additem':
        assert novelty_arg1, "For additem novelty, novelty_arg1 (name of the item to add) is needed"
        return AddItem(env, difficulty, novelty_arg1)
    if novelty_name == 'axe':
        assert novelty_arg1 in ['wooden', 'iron'], \
            "For axe novelty, novelty_arg1 (attribute of axe, e.g. wooden, iron) is needed"
        if novelty_arg2:
            assert novelty_arg2 in ['true', 'false'], \
                "For axe novelty, novelty_arg2 (breakincrease) must be 'true' or 'false'"
            return _AXE[difficulty](env, novelty_arg1, novelty_arg2)
        return _AXE[difficulty](env, novelty_arg1)
    if novelty_name == 'axetobreak':
        assert novelty_arg1 in ['wooden', 'iron'], \
            "For axe novelty, novelty_arg1 (attribute of axe, e.g. wooden, iron) is needed"
        return _AXETOBREAK[difficulty](env, novelty_arg1)
    if novelty_name == 'breakincrease':
        if novelty_arg1:
            assert novelty_arg1 in env.items, novelty_arg1 + " is not in " + env.env_id
            return BreakIncrease(env, novelty_arg1)
        return BreakIncrease(env)
    if novelty_name == 'crate':
        return Crate(env, difficulty)
    if novelty_name == 'extractincdec':
        assert novelty_arg1 in ['increase', 'decrease'], \
            "For extractincdec novelty, novelty_arg1 ('increase', 'decrease') is needed"
        assert env.env_id != 'NovelGridworld-Bow-v0', "There is nothing to extract in NovelGridworld-Bow-v0"
        if env.env_id == 'NovelGridworld-Bow-v1':
            assert novelty_arg1 == 'decrease', \
                "In NovelGridworld-Bow-v1, increasing string extraction will not benefit as only 3 string are needed"
        assert not env.env_id.startswith('NovelGridworld-Pogostick'), \
            "In NovelGridworld-Pogostick, you should not use extractincdec novelty because rubber extraction " \
            "cannot be decreased, and increasing rubber extraction will not benefit as only 1 rubber is needed"
        return ExtractIncDec(env, novelty_arg1)
    if novelty_name == 'fence':
        assert novelty_arg1, "For fence novelty, novelty_arg1 (attribute of fence, e.g. oak, jungle) is needed"
        return Fence(env, difficulty, novelty_arg1)
    if novelty_name == 'fencerestriction':
        assert novelty_arg1, \
            "For fencerestriction novelty, novelty_arg1 (attribute of fence, e.g. oak, jungle) is needed"
        return FenceRestriction(env, difficulty, novelty_arg1)
    if novelty_name == 'firewall':
        return FireWall(env, difficulty)
    if novelty_name == 'remapaction':
        return remap_action_difficulty(env, difficulty)
    assert novelty_arg1 and novelty_arg2, \
        "For replaceitem novelty, novelty_arg1 (Item to replace) and novelty_arg2(Item to replace with) are needed"
    return ReplaceItem(env, difficulty, novelty_arg1, novelty_arg2)
