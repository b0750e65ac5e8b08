"""fancyrec_b200 -- B200-native (sm_100a) implementation of FancyRec's brand x post scoring +
ranking hot path behind the reference's own Python API:

    fancyrec_b200.evaluator   <->  evaluator.py      (l2norm, cal_sim, encode_data, test_post_ranking)
    fancyrec_b200.model       <->  model.py          (l2norm, L1Penalty, BrandAspects, FancyRec shell)
    fancyrec_b200.loss        <->  loss.py           (TripletLoss, LabLoss, *_sim)
    fancyrec_b200.loss_ctrs   <->  loss_ctrs.py      (ContrastiveLoss, CrossCLR_onlyIntraModality)
    fancyrec_b200.util.ndcg / util.metric / util.imgbigfile / util.constant

All device work goes through libfrx_b200.so (C ABI: include/frx.h, sources: fancyrec_b200/csrc).
No Triton, no multi-backend dispatch, no CPU fallback.
"""
__version__ = "0.1.0"
