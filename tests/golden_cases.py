"""The reference's own golden vectors for the VPP path.

Every CRC below is a literal from the reference's test sources (file:line cited per row); the
input of every row is decoded frame #1 of tests/resources/bbb_1080x608_420_10.h264 and the CRC
is FFmpeg's av_crc(AV_CRC_32_IEEE, init=0xFFFFFFFF) over the raw output buffer.
Shared by the CPU-oracle pinning test and the GPU parity tests.
"""
# numbering == reference include/VideoProcessor.h:20-64
Y800, RGB24, BGR24, NV12, UYVY, YUV444, HSV = range(7)
PLANAR, MERGED = 0, 1
NEAREST, BILINEAR, BICUBIC, AREA = range(4)

# (name, kwargs, accepted CRCs, citation)
CRC_CASES = [
    ("nv12_passthrough", dict(fourcc=NV12), (2957341121,), "VPPTests.cpp:211-216"),
    ("y800", dict(fourcc=Y800), (3265466497,), "VPPTests.cpp:176-181"),
    ("rgb_merged", dict(fourcc=RGB24, planes=MERGED), (2225932432,), "VPPTests.cpp:134-139"),
    ("rgb_planar", dict(fourcc=RGB24, planes=PLANAR), (3151499217,), "VPPTests.cpp:141-146"),
    ("bgr_merged", dict(fourcc=BGR24, planes=MERGED), (2467105116,), "VPPTests.cpp:162-167"),
    ("bgr_planar", dict(fourcc=BGR24, planes=PLANAR), (3969775694,), "VPPTests.cpp:169-174"),
    ("rgb_540x304_nearest", dict(fourcc=RGB24, dst=(540, 304)), (3545075074,), "VPPTests.cpp:148-153"),
    ("rgb_2160x1216_nearest", dict(fourcc=RGB24, dst=(2160, 1216)), (97423732,), "VPPTests.cpp:155-160"),
    ("bgr_540x304_nearest", dict(fourcc=BGR24, dst=(540, 304)), (201454032,), "PythonTests.cpp:146-148"),
    ("uyvy", dict(fourcc=UYVY), (1323730732,), "VPPTests.cpp:183-188"),
    ("uyvy_720x480", dict(fourcc=UYVY, dst=(720, 480)), (1564587937,), "VPPTests.cpp:190-195"),
    ("yuv444", dict(fourcc=YUV444), (1110927649,), "VPPTests.cpp:197-202"),
    ("yuv444_720x480", dict(fourcc=YUV444, dst=(720, 480)), (449974214,), "VPPTests.cpp:204-209"),
    ("nv12_720x480_nearest", dict(fourcc=NV12, dst=(720, 480)), (1200915282,), "VPPTests.cpp:218-223"),
    # crop only (NV12 out)
    ("crop_left", dict(fourcc=NV12, crop=(0, 0, 320, 240)), (3435719157,), "VPPTests.cpp:225-232"),
    ("crop_center", dict(fourcc=NV12, crop=(320, 240, 720, 480)), (1515981907,), "VPPTests.cpp:234-241"),
    ("crop_center2", dict(fourcc=NV12, crop=(400, 240, 720, 480)), (655388614,), "VPPTests.cpp:243-250"),
    ("crop_right", dict(fourcc=NV12, crop=(640, 360, 1080, 608)), (602193072,), "VPPTests.cpp:252-259"),
    # crop + nearest resize (NV12 out)
    ("crop_up_left", dict(fourcc=NV12, crop=(0, 0, 320, 240), dst=(720, 480)), (1764198598,), "VPPTests.cpp:261-267"),
    ("crop_up_center", dict(fourcc=NV12, crop=(160, 120, 480, 360), dst=(720, 480)), (1834204062,), "VPPTests.cpp:269-275"),
    ("crop_up_right", dict(fourcc=NV12, crop=(400, 240, 720, 480), dst=(720, 480)), (1750083777,), "VPPTests.cpp:277-283"),
    ("crop_down_left", dict(fourcc=NV12, crop=(0, 0, 720, 480), dst=(480, 320)), (3477030875,), "VPPTests.cpp:285-291"),
    ("crop_down_right", dict(fourcc=NV12, crop=(480, 340, 1080, 608), dst=(480, 320)), (2394953726,), "VPPTests.cpp:293-299"),
    # resize-type sweep, RGB24 merged
    ("rgb_nearest_480x360", dict(dst=(480, 360), resize_type=NEAREST), (3234932936,), "PythonTests.cpp:183-185"),
    ("rgb_nearest_540x304", dict(dst=(540, 304), resize_type=NEAREST), (3545075074,), "PythonTests.cpp:187-189"),
    ("rgb_nearest_1920x1080", dict(dst=(1920, 1080), resize_type=NEAREST), (867059050,), "PythonTests.cpp:191-193"),
    ("rgb_bilinear_480x360", dict(dst=(480, 360), resize_type=BILINEAR), (1166179972,), "PythonTests.cpp:195-197"),
    ("rgb_bilinear_540x304", dict(dst=(540, 304), resize_type=BILINEAR), (2257004891,), "PythonTests.cpp:199-201"),
    ("rgb_bilinear_1920x1080", dict(dst=(1920, 1080), resize_type=BILINEAR), (930427804,), "PythonTests.cpp:203-205"),
    ("rgb_bicubic_480x360", dict(dst=(480, 360), resize_type=BICUBIC), (4261607874, 1267073424), "PythonTests.cpp:207-209"),
    ("rgb_bicubic_540x304", dict(dst=(540, 304), resize_type=BICUBIC), (4169518778,), "PythonTests.cpp:211-213"),
    ("rgb_bicubic_1920x1080", dict(dst=(1920, 1080), resize_type=BICUBIC), (2402019758,), "PythonTests.cpp:215-217"),
    ("rgb_area_480x360", dict(dst=(480, 360), resize_type=AREA), (3175240744,), "PythonTests.cpp:219-221"),
    ("rgb_area_540x304", dict(dst=(540, 304), resize_type=AREA), (2257004891,), "PythonTests.cpp:223-225"),
    ("rgb_area_1920x1080", dict(dst=(1920, 1080), resize_type=AREA), (2026855,), "PythonTests.cpp:227-229"),
    # crop + AREA, RGB24 merged
    ("crop_area_up_left", dict(dst=(1920, 1080), crop=(0, 0, 320, 240), resize_type=AREA), (2884432201,), "PythonTests.cpp:231-233"),
    ("crop_area_up_center", dict(dst=(1920, 1080), crop=(320, 240, 720, 480), resize_type=AREA), (2674082046,), "PythonTests.cpp:235-237"),
    ("crop_area_up_right", dict(dst=(1920, 1080), crop=(720, 480, 1080, 608), resize_type=AREA), (4006833449,), "PythonTests.cpp:239-241"),
    ("crop_area_down_center", dict(dst=(320, 240), crop=(120, 60, 960, 540), resize_type=AREA), (1183295093,), "PythonTests.cpp:243-245"),
]

# float32 goldens: 320x240 nearest + normalisation (VPPTests.cpp:340-512); files live in the
# reference's tests/resources/test_references/.  sha256 of each whole file is committed in
# tests/golden/float_goldens.json by make_golden.py so the GPU box (no /root/reference) can
# still check them.
FLOAT_CASES = [
    ("RGB24Normalization_320x240.yuv", dict(fourcc=RGB24, planes=MERGED, normalization=True, dst=(320, 240))),
    ("BGR24Normalization_320x240.yuv", dict(fourcc=BGR24, planes=MERGED, normalization=True, dst=(320, 240))),
    ("Y800Normalization_320x240.yuv", dict(fourcc=Y800, planes=MERGED, normalization=True, dst=(320, 240))),
    ("UYVYNormalization_320x240.yuv", dict(fourcc=UYVY, planes=MERGED, normalization=True, dst=(320, 240))),
    ("YUV444Normalization_320x240.yuv", dict(fourcc=YUV444, planes=MERGED, normalization=True, dst=(320, 240))),
    ("NV12Normalization_320x240.yuv", dict(fourcc=NV12, planes=MERGED, normalization=True, dst=(320, 240))),
    ("HSV_320x240.yuv", dict(fourcc=HSV, planes=MERGED, normalization=True, dst=(320, 240))),
]

# multi-frame goldens (CRC over the concatenated per-frame outputs), CPU only (inputs are not
# committed: they need the reference's streams)
MULTIFRAME_CASES = [
    ("bbb_1080x608_420_10.h264", 10, dict(fourcc=RGB24, dst=(720, 480)), 249831002, "WrapperTests.cpp:145"),
    ("bbb_1080x608_420_10.h264", 10, dict(fourcc=RGB24, dst=(1920, 1080)), 1775796233, "WrapperTests.cpp:240"),
    ("billiard_1920x1080_420_100.h264", 10, dict(fourcc=BGR24, dst=(720, 480)), 3048624823, "WrapperTests.cpp:241"),
    ("bbb_720x480_RGB24_250.h264", 1, dict(fourcc=RGB24, dst=(720, 480)), 2018747012, "PythonTests.cpp:277-287"),
]
