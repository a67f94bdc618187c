"""Lower a model built with the mirrored builder API (quokkasim_b200.api objects) into oracle rows.

TEST INFRASTRUCTURE (like everything under oracle/): duck-typed so that it does not import the product.
Used by tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
"""
from . import oracle as O


def lower_to_oracle(model):
    """Returns (OracleModel, {id(Distribution): [oracle distribution-instance ids]})."""
    om = O.OracleModel()
    dist_ids = {}
    comps = [c.component for c in model.components]
    index = {id(c): i for i, c in enumerate(comps)}
    for c in comps:
        kind = getattr(c, "_container_kind", None) or c.KIND
        if kind == O.DISCRETE_STOCK:
            i = om.add_component(kind, c.low_capacity, c.max_capacity, len(c.resource))
        elif kind == O.CONTAINER_STOCK:
            i = om.add_component(kind, c.low_capacity, c.max_capacity, 0)
            if len(c.resource):
                om.set_initial_containers(
                    i, [(it.resource.values if it.resource is not None else [0.0, 0.0, 0.0]) for it in c.resource],
                    [it.resource is not None for it in c.resource], [it.capacity for it in c.resource])
        elif hasattr(c, "max_process_count"):
            i = om.add_component(kind, 0, c.max_process_count, 0)
        else:
            i = om.add_component(kind)
        assert i == index[id(c)]
        if kind == O.CONTAINER_SOURCE:
            f = c.item_factory
            q = f.create_quantity_distr
            did = om.set_container_factory(i, O.make_dist(q.kind, q.stream, q.params) if q is not None else None,
                                           f.create_component_split, f.capacity)
            if q is not None:
                dist_ids.setdefault(id(q), []).append(did)
        if hasattr(c, "vector_rows"):
            dim, low, maxc, init = c.vector_rows()
            om.set_vector(i, dim, low, maxc, init, getattr(c, "total_mask", None))
            if getattr(c, "n_ports", None):
                om.set_splits(i, list(c.split_ratios))
        if hasattr(c, "process_time_distr"):
            for which, d in ((0, c.process_time_distr), (1, c.process_quantity_distr)):
                did = om.set_dist(i, which, O.make_dist(d.kind, d.stream, d.params))
                dist_ids.setdefault(id(d), []).append(did)
            for dm in c.delay_modes:
                ud, uf = dm.until_delay_distr, dm.until_fix_distr
                _, a, b = om.add_delay_mode(i, O.make_dist(ud.kind, ud.stream, ud.params),
                                            O.make_dist(uf.kind, uf.stream, uf.params))
                dist_ids.setdefault(id(ud), []).append(a)
                dist_ids.setdefault(id(uf), []).append(b)
    conns = sorted((x for c in comps for x in c._conns), key=lambda x: x[0])
    for _, a, b, n in conns:
        om.connect(index[id(a)], index[id(b)], n)
    for t, env, state in model.ext:
        om.schedule_env(t, index[id(env.component)], state)
    for t, stock, value in getattr(model, "ext_low", []):
        om.schedule_low_capacity(t, index[id(stock.component)], value)
    for t, proc, d in getattr(model, "ext_qty", []):
        did = om.schedule_process_quantity(t, index[id(proc.component)], O.make_dist(d.kind, d.stream, d.params))
        dist_ids.setdefault(id(d), []).append(did)
    return om, dist_ids
