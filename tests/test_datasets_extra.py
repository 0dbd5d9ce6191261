"""Remaining builtin datasets: exact generators (BA-Shapes family) and shape-faithful synthetic collections."""
import torch

import dgl_b200 as dgl

D = dgl.data


def test_explainability_benchmarks_are_exact():
    ds = D.BAShapeDataset(seed=1)
    g = ds[0]
    assert g.num_nodes() == 300 + 80 * 5 and ds.num_classes == 4
    assert g.ndata["label"].bincount().tolist() == [300, 160, 160, 80]
    t = D.TreeCycleDataset()[0]
    assert t.num_nodes() == 2 ** 9 - 1 + 60 * 6 and int(t.ndata["label"].sum()) == 360
    assert D.TreeGridDataset()[0].num_nodes() == 2 ** 9 - 1 + 80 * 9
    c = D.BACommunityDataset()
    assert c[0].num_nodes() == 1400 and c.num_classes == 8 and c[0].ndata["feat"].shape == (1400, 10)
    b = D.BA2MotifDataset(num_graphs=10)
    assert len(b) == 10 and b[0][0].num_nodes() == 25 and b.labels.tolist() == [0, 1] * 5


def test_aliases_and_collections():
    assert D.KarateClub is D.KarateClubDataset and D.QM9 is D.QM9Dataset and D.SST is D.SSTDataset
    assert D.AmazonCoBuy("photo")[0].num_nodes() == 7650
    assert D.CitationGraphDataset("cora")[0].num_nodes() == 2708
    p = D.PATTERNDataset(scale=0.001)
    g = p[0]
    assert "label" in g.ndata and int(g.ndata["label"].max()) <= 1
    assert D.CLUSTERDataset(mode="test", scale=0.004).num_classes == 6
    for cls in (D.MNISTSuperPixelDataset, D.CIFAR10SuperPixelDataset, D.PeptidesStructuralDataset):
        ds = cls(scale=0.0005)
        g, y = ds[0]
        assert g.num_nodes() > 1 and y is not None
    ml = D.MovieLensDataset(scale=0.05)[0]
    assert set(ml.ntypes) == {"user", "movie"} and "rate" in ml.edges["user-movie"].data
    m = ml.edges["user-movie"].data
    assert int(m["train_mask"].sum() + m["valid_mask"].sum() + m["test_mask"].sum()) == ml.num_edges("user-movie")
    q = D.QM9Edge(scale=0.001) if "scale" in D.QM9Dataset.__init__.__code__.co_varnames else D.QM9Edge()
    assert "edge_attr" in q[0][0].edata
