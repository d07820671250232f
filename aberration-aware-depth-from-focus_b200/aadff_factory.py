"""Lens / dataset factory (mirror of the reference's dff/factory.py:4-51)."""
from aadff_lens import PSFNet, ThinLens


def _make(args, split):
    cfg = args[split]
    ks, res, device = args['ks'], args['res'], args['device']
    if cfg['lens'] == 'thinlens':
        size = [float(i) for i in cfg['sensor_size']]
        return ThinLens(foc_len=cfg['foc_len'], fnum=cfg['fnum'], kernel_size=ks, sensor_size=size,
                        sensor_res=res).to(device)
    lens = PSFNet(filename=cfg['lens'], sensor_res=res, kernel_size=ks, device=device)
    lens.load_net(cfg['psfnet_path'])
    return lens


def get_lens(args):
    """args: the YAML dict of configs/aber_aware_dff_*.yml -> (train_lens, test_lens)."""
    return _make(args, 'train'), _make(args, 'test')


def get_dataset(args):
    """(train_set, test_set) exactly as dff/factory.py:33-51 selects them.  The Dataset classes themselves
    (file decoding + CPU augmentation, dff/dataset.py) are not rebuilt here: they are taken from the reference's
    ``dff.dataset``, which the shadow ``dff`` package resolves when the reference checkout is on sys.path."""
    from dff import dataset as ds
    name = args['train']['dataset']
    if name == 'Matterport3D':
        train_set = ds.Matterport3D(args['train_aif_dir'], args['train_depth_dir'], resize=args['res'])
    elif name == 'FlyingThings3D':
        train_set = ds.FlyingThings3D(args['FlyingThings3D_train'], resize=args['res'])
    else:
        raise NotImplementedError
    name = args['test']['dataset']
    if name == 'Middlebury2014':
        test_set = ds.Middlebury(args['Middlebury2014_val'], resize=args['res'], train=False)
    elif name == 'Middlebury2021':
        test_set = ds.Middlebury(args['Middlebury2021_val'], resize=args['res'], train=False)
    elif name == 'RealWorld':
        test_set = ds.RealWorld(args['RealWorld_val'], resize=args['res'], depth=False)
    else:
        raise NotImplementedError
    return train_set, test_set
