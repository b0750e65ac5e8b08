"""Drop-in for the reference's tester.py (tester.py:26-113): the same command line, checkpoint layout, output-path rule
and printed lines, with the evaluation itself (encode_data -> test_post_ranking) on the B200 kernels.

    python -m fancyrec_b200.tester insCartest --rootpath $rootpath --overwrite 1 --n_caption 1 --batch_size 64 \\
           --logger_name $logger_name                                   # bin/test_instance.sh:10

What is the reference's and stays the reference's: the data pipeline (caption files, vocabularies, bigfile datasets,
collate functions: util/data_provider.py, preprocess/) and the learned encoders (model.py VisualEncoder / Text*Encoder).
`main` obtains them through `build(opt, options, checkpoint) -> (model, data_loader)`; the default builder imports the
reference's own modules from sys.path (run it from the reference tree, as bin/test_instance.sh does) and assembles
them around THIS package's FancyRec shell, BrandAspects, MFC and projection head.  Checkpoints are the reference's:
{'epoch', 'model': [vid, text, brand, fusion state_dicts], 'best_rsum', 'opt': Namespace, 'Eiters'} (trainer.py:294-301).
"""
from __future__ import print_function

import argparse
import json
import logging
import os
import sys

import torch

from . import evaluator
from .evaluator import test_post_ranking
from .util.constant import ROOT_PATH, device


def parse_args(argv=None):
    """tester.py:26-42 -- same positional, flags, types, defaults and choices."""
    parser = argparse.ArgumentParser()
    parser.add_argument('testCollection', type=str, help='test collection')
    parser.add_argument('--rootpath', type=str, default=ROOT_PATH, help='path to datasets. (default: %s)' % ROOT_PATH)
    parser.add_argument('--overwrite', type=int, default=0, choices=[0, 1], help='overwrite existed file. (default: 0)')
    parser.add_argument('--log_step', default=10, type=int, help='Number of steps to print and record the log.')
    parser.add_argument('--batch_size', default=128, type=int, help='Size of a training mini-batch.')
    parser.add_argument('--workers', default=0, type=int, help='Number of data loader workers.')
    parser.add_argument('--logger_name', default='runs', help='Path to save the model and Tensorboard log.')
    parser.add_argument('--checkpoint_name', default='model_best.pth.tar', type=str,
                        help='name of checkpoint (default: model_best.pth.tar)')
    parser.add_argument('--n_caption', type=int, default=20,
                        help='number of captions of each image/video_frames (default: 1)')
    parser.add_argument('--level_vis', type=str, default='1+2+3', help='ablation study of visual enc')
    parser.add_argument('--level_txt', type=str, default='1+2+3', help='ablation study of text enc')
    return parser.parse_args(argv)


def makedirsforfile(filename):
    """util/common.py:7-11."""
    try:
        os.makedirs(os.path.split(filename)[0])
    except OSError:
        pass


def checkToSkip(filename, overwrite):
    """util/common.py:14-23."""
    if os.path.exists(filename):
        print("%s exists." % filename),
        if overwrite:
            print("overwrite")
            return 0
        print("skip")
        return 1
    return 0


def evaluate(options, model, data_loader, log_step=10):
    """encode_data + test_post_ranking + the four printed lines of tester.py:106-113.  Returns the 8-tuple
    (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10)."""
    brands, post_embs = evaluator.encode_data(model, data_loader, log_step, logging.info)
    ranking_metrics = test_post_ranking(options.brand_num, options.metric, model, post_embs, brands)
    print('AUC[0-1]:', ranking_metrics[2])
    print('NDCG@10[0-1]:', ranking_metrics[3])
    print('NDCG@50[0-1]:', ranking_metrics[4])
    print('recall@1:', ranking_metrics[5])
    return ranking_metrics


def build_from_reference(opt, options, checkpoint):
    """Default builder: the reference's data loaders and learned encoders (imported from sys.path, tester.py:70-104)
    around this package's FancyRec shell.  Needs the reference tree on sys.path and its dataset on disk."""
    import pickle

    import model as ref_model                       # the reference's encoders (out of scope here)
    import util.data_provider as data
    from preprocess.text2vec import get_text_encoder
    from util.imgbigfile import ImageBigFile
    from util.util import read_dict

    from . import model as frx_model

    rootpath, test_collection = opt.rootpath, opt.testCollection
    caption_files = {'test': os.path.join(rootpath, test_collection, 'TextData', '%s.caption.txt' % test_collection)}
    video_feat_path = os.path.join(rootpath, test_collection, 'FeatureData', options.video_feature)
    img_feat_path = os.path.join(rootpath, test_collection, 'FeatureData', options.img_feature)
    video_feats = {'test': ImageBigFile(video_feat_path)}
    img_feats = {'test': ImageBigFile(img_feat_path)}
    assert options.visual_feat_dim == video_feats['test'].ndims
    video2frames = {'test': read_dict(os.path.join(video_feat_path, 'video2frames.txt'))}
    bow_vocab = pickle.load(open(os.path.join(rootpath, options.trainCollection, 'TextData', 'vocabulary', 'bow',
                                              options.vocab + '.pkl'), 'rb'))
    bow2vec = get_text_encoder('bow')(bow_vocab)
    options.bow_vocab_size = len(bow_vocab)
    rnn_vocab = pickle.load(open(os.path.join(rootpath, options.trainCollection, 'TextData', 'vocabulary', 'rnn',
                                              options.vocab + '.pkl'), 'rb'))
    options.vocab_size = len(rnn_vocab)
    print("=> prepare dataloader..")
    loaders = data.get_test_data_loaders(opt, caption_files, video_feats, img_feats, rnn_vocab, bow2vec, options.text_net,
                                         opt.batch_size, opt.workers, opt.n_caption, video2frames=video2frames)
    ref = ref_model.FancyRec(options)               # builds the encoder stack the options name (model.py:538-575)
    fusion = ref.fusion_encoding
    if isinstance(fusion, ref_model.PrjHeadFusionEncoder):        # same parameters, kernels of this package in eval mode
        ours = frx_model.PrjHeadFusionEncoder(options)
        ours.load_state_dict(fusion.state_dict())
        fusion = ours
    model = frx_model.FancyRec(options, vid_encoding=ref.vid_encoding, text_encoding=ref.text_encoding,
                               fusion_encoding=fusion)
    return model, loaders['test']


def main(argv=None, build=None):
    """tester.py:51-113.  `build(opt, options, checkpoint) -> (model, data_loader)` supplies the encoders and the data
    (default: build_from_reference)."""
    opt = parse_args(argv)
    print(json.dumps(vars(opt), indent=2))

    testCollection = opt.testCollection
    resume = os.path.join(opt.logger_name, opt.checkpoint_name)
    if not os.path.exists(resume):
        logging.info(resume + ' not exists.')
        sys.exit(0)

    checkpoint = torch.load(resume, map_location='cpu', weights_only=False)     # the pickled argparse Namespace rides along
    print("=> loaded!")
    options = checkpoint['opt']
    if not hasattr(options, 'concate'):
        setattr(options, "concate", "full")

    trainCollection = options.trainCollection
    output_dir = resume.replace(trainCollection, testCollection)
    output_dir = output_dir.replace('/%s/' % options.cv_name, '/results/%s/' % trainCollection)
    pred_error_matrix_file = os.path.join(output_dir, 'pred_errors_matrix.pth.tar')
    if checkToSkip(pred_error_matrix_file, opt.overwrite):
        sys.exit(0)
    makedirsforfile(pred_error_matrix_file)

    model, data_loader = (build or build_from_reference)(opt, options, checkpoint)
    model = model.to(device)
    model.load_state_dict(checkpoint['model'])
    model.Eiters = checkpoint['Eiters']
    return evaluate(options, model, data_loader, opt.log_step)


if __name__ == '__main__':
    main()
