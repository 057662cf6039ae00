"""networkx <-> Edge list adaptors (mac/utils/conversions.py:9-49).  Input adaptor only."""
from .graphs import Edge


def nx_to_mac(G):
    """conversions.py:9-31: i < j normalisation, default weight 1.0."""
    edges = []
    for i, j, data in G.edges(data=True):
        weight = data.get("weight", 1.0)
        edges.append(Edge(i, j, weight) if i < j else Edge(j, i, weight))
    return edges


def mac_to_nx(edges):
    """conversions.py:34-49."""
    import networkx as nx
    G = nx.Graph()
    for edge in edges:
        a, b = (edge.i, edge.j) if edge.i < edge.j else (edge.j, edge.i)
        G.add_edge(a, b, weight=edge.weight)
    return G
