// swift-tools-version:5.5
// Swift-on-Linux host package: the reference's SyllableDetectorConfig / SyllableDetector / TrackDetector API over the
// C-ABI of libsyldet_cuda.so.  NOT built in this repository's CI (no Swift toolchain in the image); see INTEGRATION.md.
//   swift build -Xcc -I../include -Xlinker -L../syllable-detector-swift_b200 -Xlinker -rpath -Xlinker ../syllable-detector-swift_b200
import PackageDescription

let package = Package(
    name: "SyllableDetector",
    products: [.library(name: "SyllableDetector", targets: ["SyllableDetector"])],
    targets: [
        .systemLibrary(name: "CSyldet", path: "Sources/CSyldet"),
        .target(name: "SyllableDetector", dependencies: ["CSyldet"]),
    ]
)
