"""Host-side render of ONE env from exported state (never on the GPU path).

`render_spec` lists everything the reference's render() draws (pogostick_v1_env.py:556-620): the grid image, the facing
arrow, the axis labels, the info panel (steps / facing / last action / selected item / reward / step cost / done), the
win banner and the inventory legend with its colour fractions.  `draw` hands the spec to matplotlib when it is
installed; `to_text` is the 'ansi' rendering of the same spec."""

FACING = ('NORTH', 'SOUTH', 'WEST', 'EAST')                      # pogostick_v1_env.py:33
ARROW = {'NORTH': (0, -0.01), 'SOUTH': (0, 0.01), 'WEST': (-0.01, 0), 'EAST': (0.01, 0)}


def render_spec(env_id, grid, agent_location, agent_facing_str, items_id, inventory_items_quantity, goal_item_to_craft,
                selected_item='', step_count=0, last_action='Forward', last_reward=0, last_step_cost=0, last_done=False,
                title=None):
    map_size = len(grid)
    r, c = agent_location
    x2, y2 = ARROW[agent_facing_str]
    info = '\n'.join(["               Info:             ",
                      "Steps: " + str(step_count),
                      "Agent Facing: " + agent_facing_str,
                      "Action: " + last_action,
                      "Selected item: " + selected_item,
                      "Reward: " + str(last_reward),
                      "Step Cost: " + str(last_step_cost),
                      "Done: " + str(last_done)])
    texts = [(map_size, map_size // 2, 'EAST'), (-(map_size // 2) - 0.5, 2.25, info)]
    if last_done:
        if inventory_items_quantity[goal_item_to_craft] >= 1:
            banner = "YOU WIN " + env_id + "!!!" + "\nYOU CRAFTED " + goal_item_to_craft.upper() + "!!!"
        else:
            banner = "YOU CAN'T WIN " + env_id + "!!!"
        texts.append((0 - 0.1, map_size // 2, banner))
    legend = [('agent', None), ('INVENTORY:', None)]
    for item in sorted(inventory_items_quantity):
        legend.append((item + ': ' + str(inventory_items_quantity[item]), round(items_id[item] / len(items_id), 9)))
    return {'title': env_id if title is None else title, 'grid': [[int(v) for v in row] for row in grid],
            'vmax': len(items_id), 'arrow': (c, r, x2, y2), 'axis': ('NORTH', 'SOUTH', 'WEST'), 'texts': texts,
            'legend': legend, 'facing': agent_facing_str}


def to_text(spec):
    """'ansi' mode: the grid (item ids in hex, '.' = air, the agent as ^ v < >) followed by the info panel, the banner
    and the legend lines."""
    c, r = spec['arrow'][0], spec['arrow'][1]
    mark = {'NORTH': '^', 'SOUTH': 'v', 'WEST': '<', 'EAST': '>'}[spec['facing']]
    rows = []
    for i, row in enumerate(spec['grid']):
        rows.append(' '.join(mark if (i, j) == (r, c) else ('.' if v == 0 else '%x' % v) for j, v in enumerate(row)))
    parts = ['\n'.join(rows)] + [t[2] for t in spec['texts'][1:]] + ['\n'.join(label for label, _ in spec['legend'])]
    return '\n\n'.join(parts)


def draw(spec):
    """matplotlib drawing of the spec, call for call what the reference does."""
    import matplotlib.pyplot as plt
    from matplotlib.cm import get_cmap
    from matplotlib.lines import Line2D
    color_map = "gist_ncar"
    plt.figure(spec['title'], figsize=(9, 5))
    plt.imshow(spec['grid'], cmap=color_map, vmin=0, vmax=spec['vmax'])
    c, r, x2, y2 = spec['arrow']
    plt.arrow(c, r, x2, y2, head_width=0.7, head_length=0.7, color='white')
    plt.title(spec['axis'][0], fontsize=10)
    plt.xlabel(spec['axis'][1])
    plt.ylabel(spec['axis'][2])
    x, y, s = spec['texts'][0]
    plt.text(x, y, s, rotation=90)
    x, y, s = spec['texts'][1]
    plt.text(x, y, s, fontsize=10, bbox=dict(boxstyle='round', facecolor='w', alpha=0.2))
    for x, y, s in spec['texts'][2:]:
        plt.text(x, y, s, fontsize=18, bbox=dict(boxstyle='round', facecolor='w', alpha=1))
    cmap = get_cmap(color_map)
    handles = [Line2D([0], [0], marker="^", color='w', label='agent', markerfacecolor='w', markersize=12,
                      markeredgewidth=2, markeredgecolor='k'),
               Line2D([0], [0], color='w', label="INVENTORY:")]
    for label, fraction in spec['legend'][2:]:
        handles.append(Line2D([0], [0], marker="s", color='w', label=label, markerfacecolor=cmap(fraction), markersize=16))
    plt.legend(handles=handles, bbox_to_anchor=(1.55, 1.02))
    plt.tight_layout()
    plt.pause(0.01)
    plt.clf()
